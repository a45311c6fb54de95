"""Worker of tests/test_gpu_prove.py::test_prelaunch_*: proves the m = 7 point-mult instance twice under the environment the
parent set (VPIN_PRELAUNCH_Q, VPIN_TEST_DROP_POST, VPIN_MAILBOX_TIMEOUT_MS are read once per process) and prints the digests."""
import hashlib
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from vpin_b200 import api, workloads as W


def main():
    ctx = api.Context(0)
    weights, mx, my = W.synth_point_mult(7)
    dims, inst, vp, vi, v, inputs = api.point_mult(ctx, weights, mx, my)
    sq, sp = W.tape_seeds()
    out = []
    for _ in range(2):
        t0 = time.time()
        o = api.prove_flow(ctx, dims, inst, vp, vi, v, inputs, sq, sp)
        out.append({"proof": hashlib.sha256(o["proof"]).hexdigest(), "comm": hashlib.sha256(o["comm"]).hexdigest(), "s": time.time() - t0})
        del o
    del inst
    ctx.close()
    print("RESULT " + json.dumps(out))


if __name__ == "__main__":
    main()
