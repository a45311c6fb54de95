"""Wall-clock breakdown of bench.py's resident step, call by call: python scripts/time_step.py [workload]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from vpin_b200 import api

tag = sys.argv[1] if len(sys.argv) > 1 else "A"
ctx = api.Context(0)
wl = bench.make_workload(tag)
states = []
if wl["add"] is not None:
    states.append(bench.InstanceState(ctx, "point_add", api.point_addition(ctx, *wl["add"]), torch))
states.append(bench.InstanceState(ctx, "point_mult", api.point_mult(ctx, *wl["mult"]), torch))
sq, sp = wl["seeds"]
for rep in range(3):
    for s in states:
        T = [time.time()]
        def mark():
            ctx.sync(); T.append(time.time())
        comm, decomm = api.SNARK.encode(s.inst, s.gens); mark()
        tape = api.RandomTape(b"\x02", sq)
        api.dev_poly_commit(ctx, s.gens, s.d_assign[0], s.n, tape, s.d_pts[0], s.d_blinds[0]); mark()
        api.dev_poly_commit(ctx, s.gens, s.d_assign[1], s.n, tape, s.d_pts[1], s.d_blinds[1]); mark()
        api.dev_poly_commit_with_blinds(ctx, s.gens, s.d_assign[2], s.n, s.d_blinds[0], s.d_blinds[1], s.d_pts[2], s.d_blinds[2]); mark()
        api.dev_commitments_add(ctx, s.d_pts[0], s.d_pts[1], s.gens.L, s.d_pts[3]); mark()
        wit = api.DeviceWitness(ctx, s.gens, s.d_assign[2], s.n, s.d_pts[3], s.d_blinds[2]); mark()
        proof = api.my_lib_prove_resident(s.inst, decomm, wit, s.inputs, s.gens, bench.TRANSCRIPT_LABEL, sp); mark()
        names = ["encode", "commit_para", "commit_input", "commit_vars", "comm_add", "witness", "prove"]
        if rep == 2:
            print(s.kind, "  ".join(f"{n} {1e3*(b-a):.2f}" for n, a, b in zip(names, T, T[1:])), f" total {1e3*(T[-1]-T[0]):.1f} ms")
            print("   ", {k: round(v, 2) for k, v in ctx.phase_times().items()})
        del decomm, wit
