cd $GRAFT_REPO_ROOT
N=$1
if [ "$N" = "2" ]; then timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2c_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2c_pytest.log; tail -3 gpurun_out/r2c_pytest.log; fi
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r2c_bench_${N}gpu.json 2> gpurun_out/r2c_bench_${N}gpu.err; echo "bench rc=$?"
python - <<PY
import json
d=json.load(open("gpurun_out/r2c_bench_${N}gpu.json"))
print("N=$N step", d["ms_per_step"], "e2e", d["e2e"]["value"], d["sharded_equals_unsharded"], "rows_only", d["one_proof_rows_only"]["value"], "replicas", d["replicas"].get("value"), d["replicas"].get("networks_per_s"))
print({k:(v.get("value"),v.get("matches_golden"),v.get("snark_prove_ms_point_mult")) for k,v in d["other_configs"].items()})
print(d["msm"])
PY
tail -2 gpurun_out/r2c_bench_${N}gpu.err
