cd $GRAFT_REPO_ROOT
for ov in 1 2 1 2; do
VPIN_BENCH_OVERLAP_ENCODE=$ov VPIN_BENCH_CONCURRENT=1 VPIN_BENCH_OTHER=conv3,E timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-profile > gpurun_out/r2z_bench_ov$ov.json 2> gpurun_out/r2z_bench_ov$ov.err
python - <<PY
import json
d=json.load(open("gpurun_out/r2z_bench_ov$ov.json"))
print("overlap=$ov step", d["ms_per_step"], "e2e", d["e2e"]["value"], d["parity"]["matches_golden"], {k:(v.get("value"),v.get("matches_golden")) for k,v in d["other_configs"].items()})
print(d["e2e"]["slowest_step_calls_ms"], "prove", round(d["phases_ms_point_mult"]["SNARK::prove"],2))
PY
tail -2 gpurun_out/r2z_bench_ov$ov.err
done
