"""Ad-hoc timing of the full driver flow on the GPU (not the bench): python scripts/time_flow.py <m_mult> [n_add]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from vpin_b200 import api, workloads as W

m = int(sys.argv[1])
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
ctx = api.Context(0)
print("imad peak (mad.wide.u32/s): %.3e" % ctx.imad_peak())
t = time.time(); weights, px, py = W.synth_point_mult(m); print("synth %.2fs" % (time.time() - t))
t = time.time(); dims, inst, vp, vi, v, inputs = api.point_mult(ctx, weights, px, py); print("build+Instance::new %.2fs" % (time.time() - t), dims)
sq, sp = W.tape_seeds()
for rep in range(reps):
    l0 = ctx.kernel_launches
    t0 = time.time()
    gens = api.SNARKGens(ctx, *dims); ctx.sync(); t1 = time.time()
    comm, decomm = api.SNARK.encode(inst, gens); t2 = time.time()
    tape = api.RandomTape(b"\x02", sq)
    p_para, p_input, p_vars = inst.pad(vp), inst.pad(vi), inst.pad(v)
    c_para, b_para = api.dense_mlpoly_commit(ctx, gens, p_para, tape)
    c_input, b_input = api.dense_mlpoly_commit(ctx, gens, p_input, tape)
    c_vars, b_vars = api.my_dense_mlpoly_commit(ctx, gens, p_vars, b_para, b_input)
    combined = ctx.commitments_add(c_para, c_input); t3 = time.time()
    proof = api.my_lib_prove(inst, decomm, p_vars, inputs, gens, b"snark_example", combined, b_vars, sp); t4 = time.time()
    print(f"rep {rep}: gens {t1-t0:.3f}s encode {t2-t1:.3f}s commits {t3-t2:.3f}s prove {t4-t3:.3f}s total {t4-t0:.3f}s launches {ctx.kernel_launches-l0} proof {len(proof)}B")
    for k, val in ctx.phase_times().items():
        print(f"   {k:32s} {val:10.2f} ms")
    del decomm, gens
