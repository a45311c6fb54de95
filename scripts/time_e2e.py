"""Breakdown of the host-buffer (e2e) leg: python scripts/time_e2e.py [m]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from vpin_b200 import api, workloads as W
m = int(sys.argv[1]) if len(sys.argv) > 1 else 178
ctx = api.Context(0)
dims, inst0, vp, vi, v, inputs = api.point_mult(ctx, *W.synth_point_mult(m))
A, B, Cm = inst0.export_coo(dims[1])
sq, sp = W.tape_seeds()
for rep in range(3):
    t = [time.time()]
    inst = api.Instance(ctx, dims[0], dims[1], dims[2], A, B, Cm); t.append(time.time())
    gens = api.SNARKGens(ctx, *dims); t.append(time.time())
    comm, decomm = api.SNARK.encode(inst, gens); t.append(time.time())
    tape = api.RandomTape(b"\x02", sq)
    p_para, p_input, p_vars = inst.pad(vp), inst.pad(vi), inst.pad(v); t.append(time.time())
    c_para, b_para = api.dense_mlpoly_commit(ctx, gens, p_para, tape)
    c_input, b_input = api.dense_mlpoly_commit(ctx, gens, p_input, tape)
    c_vars, b_vars = api.my_dense_mlpoly_commit(ctx, gens, p_vars, b_para, b_input); t.append(time.time())
    combined = ctx.commitments_add(c_para, c_input); t.append(time.time())
    proof = api.my_lib_prove(inst, decomm, p_vars, inputs, gens, b"snark_example", combined, b_vars, sp); t.append(time.time())
    names = ["Instance::new", "SNARKGens::new", "encode", "pad(py)", "3 commits", "comm add", "my_lib_prove"]
    print(f"rep {rep}: " + "  ".join(f"{n} {1e3*(b-a):.1f}ms" for n, a, b in zip(names, t, t[1:])) + f"  total {1e3*(t[-1]-t[0]):.1f}ms")
    print("   phases:", {k: round(x, 1) for k, x in ctx.phase_times().items()})
