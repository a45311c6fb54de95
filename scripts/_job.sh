cd $GRAFT_REPO_ROOT
timeout 1500 python bench.py --steps 20 --warmup 5 > gpurun_out/r2b_bench_1gpu.json 2> gpurun_out/r2b_bench_1gpu.err; echo "bench rc=$?"
tail -2 gpurun_out/r2b_bench_1gpu.err
bash scripts/ncu_round2b.sh > gpurun_out/r2b_ncu.log 2>&1; tail -3 gpurun_out/r2b_ncu.log
bash scripts/sanitize.sh > gpurun_out/r2b_sanitize.log 2>&1; cat gpurun_out/r2b_sanitize.log | tail -30
for tag in E L5; do timeout 600 python scripts/prove_shape_resident.py $tag 3 > gpurun_out/r2b_resident_${tag}_1gpu.log 2>&1; tail -4 gpurun_out/r2b_resident_${tag}_1gpu.log; done
