cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2y_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2y_pytest.log
tail -3 gpurun_out/r2y_pytest.log
VPIN_BENCH_CONCURRENT=1 VPIN_BENCH_OTHER=conv3,conv5,conv7,E timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2y_bench.json 2> gpurun_out/r2y_bench.err
python - <<PY
import json
d=json.load(open("gpurun_out/r2y_bench.json"))
print("step", d["ms_per_step"], "e2e", d["e2e"]["value"], d["parity"]["matches_golden"], {k:(v.get("value"),v.get("matches_golden")) for k,v in d["other_configs"].items()})
print(d["e2e"]["slowest_step_calls_ms"])
print({k:round(v,2) for k,v in d["phases_ms_point_mult"].items()})
PY
tail -3 gpurun_out/r2y_bench.err
