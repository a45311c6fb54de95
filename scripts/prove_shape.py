"""Prove one named vPIN shape on the GPU and check the proof with the oracle's verifier: python scripts/prove_shape.py E"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_lib as O
from vpin_b200 import api, workloads as W
tag = sys.argv[1] if len(sys.argv) > 1 else "E"
m, n_add = W.SHAPES[tag]
ctx = api.Context(0)
t = time.time(); weights, px, py = W.synth_point_mult(m); dims, inst, vp, vi, v, inputs = api.point_mult(ctx, weights, px, py)
print(f"{tag}: m={m} dims={dims} build {time.time()-t:.2f}s")
sq, sp = W.tape_seeds()
import torch
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
first = None
for rep in range(reps):
    t = time.time(); got = api.prove_flow(ctx, dims, inst, vp, vi, v, inputs, sq, sp); dt = time.time() - t
    for k in ("decomm", "gens"):  # drop the device-side handles now (the decommitment alone is 22 GB at L5)
        got.pop(k)
    print(f"  prove_flow (gens+encode+commits+prove, host buffers) {dt:.3f}s  proof {len(got['proof'])} B", flush=True)
    print("   ", {k: round(x, 1) for k, x in ctx.phase_times().items() if not k.startswith('batched')})
    print("    HBM in use after the flow: %.1f GB" % ((torch.cuda.mem_get_info()[1] - torch.cuda.mem_get_info()[0]) / 1e9), flush=True)
    if first is None:
        first = got
        t = time.time()
        ok = O.verify(dims, got["proof"], got["comm"], inputs, got["comm_vars_para"], got["comm_vars_input"])
        print(f"  oracle my_lib_verify -> {ok} in {time.time()-t:.2f}s", flush=True)
        assert ok == 1
    else:
        assert got["proof"] == first["proof"] and got["comm"] == first["comm"], "not deterministic"
