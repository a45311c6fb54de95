"""Calibration of bench.py's reference-arm scale (REF_SCALE): the CPU port on the FULL point-mult instance of a workload and on
the m = 18 sample, same box, same thread count:  python scripts/calibrate_reference.py [tag]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_lib as O
from vpin_b200 import workloads as W

tag = sys.argv[1] if len(sys.argv) > 1 else "A"
threads = os.cpu_count() or 1
sq, sp = W.tape_seeds()
keys = ("gens", "SNARK::encode", "witness_commits", "SNARK::prove")
res = {}
for m in (18, W.SHAPES[tag][0]):
    t0 = time.time()
    f = O.Flow(O.build_point_mult(*W.synth_point_mult(m)), sq, sp, verify=False, threads=threads)
    res[m] = sum(f.times[k] for k in keys) / 1e3
    print(f"m={m}: {res[m]:.2f} s of prover time ({time.time() - t0:.1f} s wall), phases {({k: round(f.times[k] / 1e3, 2) for k in keys})}", flush=True)
full = W.SHAPES[tag][0]
print(f"{tag}: cores={threads}  full/sample = {res[full] / res[18]:.2f}  (bench.py REF_SCALE uses 6.67 for A)")
