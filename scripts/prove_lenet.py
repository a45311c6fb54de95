"""LeNet full model on ONE B200: every R1CS instance of the seven layers (src/LeNet/Server.py:690-698, :753-761 — 7508 point
multiplications and 16864 point additions in 12 instances; layer 5's point-mult instance is the reference's 230 GB one) is
built, encoded, committed and proved through the host-buffer API, and every proof is checked with the oracle's restatement of
my_lib_verify:  python scripts/prove_lenet.py [reps]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_lib as O
from vpin_b200 import api, workloads as W

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 1
ctx = api.Context(0)
sq, sp = W.tape_seeds()
# the largest instance first: its generator tables (and window widths) then serve every smaller one
order = sorted(W.LENET_LAYERS, key=lambda t: -W.SHAPES[t][0])
for rep in range(reps):
    total_build = total_flow = 0.0
    for tag in order:
        m, n_add = W.SHAPES[tag]
        for kind, count in (("point_mult", m), ("point_add", n_add)):
            if not count:
                continue
            t = time.time()
            if kind == "point_mult":
                dims, inst, vp, vi, v, inputs = api.point_mult(ctx, *W.synth_point_mult(count, seed=W.SEED + int(tag[1:])))
            else:
                dims, inst, vp, vi, v, inputs = api.point_addition(ctx, *W.synth_point_add(count, seed=W.SEED + 1 + int(tag[1:]), infinity_every=97))
            tb = time.time() - t
            t = time.time()
            got = api.prove_flow(ctx, dims, inst, vp, vi, v, inputs, sq, sp)
            tf = time.time() - t
            prove_ms = ctx.phase_times().get("SNARK::prove", 0.0)
            for k in ("decomm", "gens"):
                got.pop(k)
            ok = O.verify(dims, got["proof"], got["comm"], inputs, got["comm_vars_para"], got["comm_vars_input"]) if rep == 0 else 1
            assert ok == 1, f"{tag} {kind}: proof rejected"
            total_build += tb; total_flow += tf
            print(f"{tag} {kind:10s} n={count:5d} cons={dims[0]:9d}  build {tb:6.2f}s  gens+encode+commits+prove {tf:6.3f}s (SNARK::prove {prove_ms:7.1f} ms)  "
                  f"proof {len(got['proof'])} B  verified={ok == 1}", flush=True)
            del inst, got
    print(f"rep {rep}: LeNet total: build {total_build:.2f}s + prove flows {total_flow:.2f}s = {total_build + total_flow:.2f}s", flush=True)
