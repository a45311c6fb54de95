"""GPU, two ranks: the product's multi-GPU code (NCCL all-gather of the sharded Hyrax rows, instance-sharded batched sumcheck
rounds) must give the single-GPU bytes. Needs two GPUs (`gpurun --gpus 2`); the driver's one-GPU test box skips it - there the
same check runs inside bench.py at N = 2, 4, 8 (`sharded_equals_unsharded`, `one_proof_rows_only.matches_golden`)."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_one_proof_on_two_gpus_matches_the_golden_digests():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    p = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", "29517", os.path.join(ROOT, "tests", "dist_worker.py"), "conv3", "A"],
                       capture_output=True, text=True, timeout=900, cwd=ROOT, env=env)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-4000:]
    line = [l for l in p.stdout.splitlines() if l.startswith("DIST_RESULT")]
    assert line, p.stdout[-2000:]
    res = json.loads(line[0].split(" ", 1)[1])
    assert res["world"] == 2 and res["cases"] == 8 and res["bad"] == [], res
