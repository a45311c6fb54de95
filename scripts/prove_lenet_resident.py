"""LeNet full model on ONE B200 with the device-resident driver flow: the seven point-mult instances (7508 multiplications; layer
5's is the reference's 230 GB / ~4 h one) are expanded, encoded, committed and proved without a large buffer crossing PCIe
(api.prove_flow_resident), the five small point-add instances go through the host-buffer flow; every proof of the first pass is
checked with the oracle's restatement of my_lib_verify.   python scripts/prove_lenet_resident.py [passes]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_lib as O
from vpin_b200 import api, workloads as W

passes = int(sys.argv[1]) if len(sys.argv) > 1 else 2
ctx = api.Context(0)
sq, sp = W.tape_seeds()
# the largest instance first: its generator tables (and window geometry) then serve every smaller one
order = sorted(W.LENET_LAYERS, key=lambda t: -W.SHAPES[t][0])
inputs_of = {}
for tag in order:
    m, n_add = W.SHAPES[tag]
    if m:
        inputs_of[(tag, "point_mult")] = W.synth_point_mult(m, seed=W.SEED + int(tag[1:]))
    if n_add:
        inputs_of[(tag, "point_add")] = W.synth_point_add(n_add, seed=W.SEED + 1 + int(tag[1:]), infinity_every=97)
for p in range(passes):
    total = 0.0
    for (tag, kind), data in inputs_of.items():
        t = time.time()
        if kind == "point_mult":
            got = api.prove_flow_resident(ctx, *data, sq, sp)
            dims, inputs = got["dims"], got["inputs"]
        else:
            dims, inst, vp, vi, v, inputs = api.point_addition(ctx, *data)
            got = api.prove_flow(ctx, dims, inst, vp, vi, v, inputs, sq, sp)
            for k in ("decomm", "gens"):
                got.pop(k)
            del inst
        ctx.sync()
        dt = time.time() - t
        total += dt
        prove_ms = ctx.phase_times().get("SNARK::prove", 0.0)
        ok = O.verify(dims, got["proof"], got["comm"], inputs, got["comm_vars_para"], got["comm_vars_input"]) if p == 0 else 1
        assert ok == 1, f"{tag} {kind}: proof rejected"
        print(f"pass {p} {tag} {kind:10s} cons={dims[0]:9d}  whole flow {dt:6.3f} s (SNARK::prove {prove_ms:7.1f} ms)  proof {len(got['proof'])} B"
              + ("  verified" if p == 0 else ""), flush=True)
        del got
    print(f"pass {p} ({'cold: generator tables built' if p == 0 else 'warm'}): LeNet, all {len(inputs_of)} instances: {total:.2f} s", flush=True)
