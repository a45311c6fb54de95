"""Per-kernel-class device times of ONE prove with every scope timed (min_units = 0): python scripts/prof_small.py [m]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from vpin_b200 import api, workloads as W
m = int(sys.argv[1]) if len(sys.argv) > 1 else 178
ctx = api.Context(0)
dims, inst, vp, vi, v, inputs = api.point_mult(ctx, *W.synth_point_mult(m))
sq, sp = W.tape_seeds()
gens = api.SNARKGens(ctx, *dims)
comm, decomm = api.SNARK.encode(inst, gens)
tape = api.RandomTape(b"\x02", sq)
p_para, p_input, p_vars = inst.pad(vp), inst.pad(vi), inst.pad(v)
c_para, b_para = api.dense_mlpoly_commit(ctx, gens, p_para, tape)
c_input, b_input = api.dense_mlpoly_commit(ctx, gens, p_input, tape)
c_vars, b_vars = api.my_dense_mlpoly_commit(ctx, gens, p_vars, b_para, b_input)
combined = ctx.commitments_add(c_para, c_input)
for rep in range(2):
    proof = api.my_lib_prove(inst, decomm, p_vars, inputs, gens, b"snark_example", combined, b_vars, sp)
ctx.profile_enable(True, 0.0)
t0 = time.time()
proof = api.my_lib_prove(inst, decomm, p_vars, inputs, gens, b"snark_example", combined, b_vars, sp)
t1 = time.time()
prof, madds = ctx.profile_read()
print(f"prove with all scopes timed: {1e3*(t1-t0):.1f} ms")
for k, v in sorted(prof.items(), key=lambda kv: -kv[1]["ms"]):
    if v["launches"]:
        print(f"  {k:24s} scopes-launches {v['launches']:5d}  total {v['ms']:8.3f} ms  avg/launch {1e3*v['ms']/v['launches']:8.1f} us")
