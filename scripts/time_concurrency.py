"""What the concurrent point-add proof costs the point-mult proof (the step's critical path), with and without the urgent
stream priority: python scripts/time_concurrency.py [workload]"""
import argparse, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench

tag = sys.argv[1] if len(sys.argv) > 1 else "A"
args = argparse.Namespace(steps=5, warmup=3)
for name, prio, with_add in (("point-mult alone", "0", False), ("both, equal priority", "0", True), ("both, point-mult urgent", "1", True)):
    os.environ["VPIN_BENCH_PRIORITY"] = prio
    wl = bench.make_workload(tag)
    if not with_add:
        wl["add"] = None
    leg = bench.Leg(args, torch, None, wl, distributed=False)
    r = leg.time_resident(sample_clocks=False)
    print(f"{name:28s} {1e3 * r['step_s']:.2f} ms/step   SNARK::prove(point-mult) {r['phases'].get('SNARK::prove', 0):.1f} ms", flush=True)
    leg.close()
