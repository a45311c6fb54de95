cd $GRAFT_REPO_ROOT
N=$1
if [ "$N" = "1" ]; then
timeout 1500 python bench.py --steps 20 --warmup 5 > gpurun_out/r2g_bench_1gpu.json 2> gpurun_out/r2g_bench_1gpu.err; echo "bench rc=$?"
else
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r2g_bench_${N}gpu.json 2> gpurun_out/r2g_bench_${N}gpu.err; echo "bench rc=$?"
fi
python - <<PY
import json
d=json.load(open("gpurun_out/r2g_bench_${N}gpu.json"))
print("N=$N step", d["ms_per_step"], "e2e", d["e2e"]["value"], d["sharded_equals_unsharded"], d["parity"]["matches_golden"])
print({k:(v.get("value"),v.get("matches_golden"),v.get("snark_prove_ms_point_mult")) for k,v in d["other_configs"].items()})
print(d["msm"]["mpoints_per_s"], d.get("replicas") and d["replicas"].get("value"))
PY
tail -2 gpurun_out/r2g_bench_${N}gpu.err
