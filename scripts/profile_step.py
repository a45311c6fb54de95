"""One warm bench step (both CNN-A instances, concurrently as in bench.py) between cudaProfilerStart/Stop, for
  ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/step.csv \
      python scripts/profile_step.py [workload]
so that the launch list holds exactly the kernels of one timed step."""
import argparse, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench

tag = sys.argv[1] if len(sys.argv) > 1 else "A"
args = argparse.Namespace(steps=1, warmup=2)
leg = bench.Leg(args, torch, None, bench.make_workload(tag), distributed=False)
for _ in range(2):
    leg.one_step_resident()
leg.sync_all()
torch.cuda.profiler.start()
leg.one_step_resident()
leg.sync_all()
torch.cuda.profiler.stop()
print("launches so far:", leg.ctx.kernel_launches)
leg.close()
