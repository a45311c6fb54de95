"""GPU: the run-time alternatives of the MSM finishing kernels and of the table geometry must give the oracle's bytes too.
They are chosen by environment variables that the library reads once per process, so each combination runs the MSM / Hyrax
kernel tests and one whole proof (point-mult, m = 7) in a pytest subprocess:
  VPIN_MSM_SUB / VPIN_MSM_W   two sub-tables per generator (the Horner pass twice as long), every window width
  VPIN_SEGSUM_LANES           lanes per (row, window) pair in the segment sum: serial, and the full 32-lane tree
  VPIN_TREE_QUAD=0            one lane per addition in the shuffle trees (the shipped trees use four)
  VPIN_HORNER_QUAD=0          one thread per row in the window Horner pass (the shipped pass uses four lanes)
  VPIN_PRELAUNCH_Q=0          challenges as kernel parameters (no round is launched ahead of its challenge)"""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

VARIANTS = [
    {"VPIN_MSM_SUB": "2", "VPIN_MSM_W": "13"},
    {"VPIN_MSM_SUB": "2", "VPIN_MSM_W": "15"},
    {"VPIN_MSM_SUB": "4", "VPIN_MSM_W": "12"},
    {"VPIN_MSM_W": "14"},
    {"VPIN_SEGSUM_LANES": "1"},
    {"VPIN_SEGSUM_LANES": "32"},
    {"VPIN_TREE_QUAD": "0", "VPIN_HORNER_QUAD": "0"},
    {"VPIN_PRELAUNCH_Q": "0", "VPIN_DEREFS_EARLY": "0"},
]


@pytest.mark.parametrize("env", VARIANTS, ids=lambda e: ",".join(f"{k[5:]}={v}" for k, v in e.items()))
def test_variant_gives_the_oracle_bytes(env):
    e = dict(os.environ)
    e.update(env)
    sel = "msm_matches_oracle or msm_small_and_negative or hyrax_commit_matches_oracle or point_mult_flow_m7"
    p = subprocess.run([sys.executable, "-m", "pytest", os.path.join(ROOT, "tests", "test_gpu_kernels.py"),
                        os.path.join(ROOT, "tests", "test_gpu_prove.py"), "-m", "gpu", "-x", "-q", "-k", sel],
                       env=e, capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert p.returncode == 0, p.stdout[-3000:] + p.stderr[-2000:]
    assert " passed" in p.stdout and "failed" not in p.stdout, p.stdout[-1000:]
