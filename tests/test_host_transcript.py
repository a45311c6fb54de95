"""Host logic, no GPU: the product's Merlin/STROBE/Keccak (vpin_b200/csrc/merlin.hpp) against the oracle's restatement, and the
portable paths of the limb routines (limbs.cuh) against the routines they specialise."""
import os
import subprocess
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_product_transcript_equals_the_oracle_transcript():
    src = os.path.join(ROOT, "tests", "cpp", "transcript_parity.cpp")
    with tempfile.TemporaryDirectory() as d:
        exe = os.path.join(d, "transcript_parity")
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-o", exe, src])
        out = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and "mismatches=0" in out.stdout, out.stdout + out.stderr


def _run_cpp(name):
    src = os.path.join(ROOT, "tests", "cpp", name + ".cpp")
    with tempfile.TemporaryDirectory() as d:
        exe = os.path.join(d, name)
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-o", exe, src])
        return subprocess.run([exe], capture_output=True, text=True, timeout=300)


def test_squaring_and_montgomery_reduction_equal_the_general_routines():
    out = _run_cpp("limbs_host")
    assert out.returncode == 0 and "mismatches=0" in out.stdout, out.stdout + out.stderr
