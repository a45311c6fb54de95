mkdir -p gpurun_out
./scripts/ubench/imad_rates2 > gpurun_out/r2b_imad_rates2.log 2>&1; cat gpurun_out/r2b_imad_rates2.log
for v in 0 1 2 3 4 5 11 12 13 14 15 16; do echo "== variant $v"; VPIN_MSM_VARIANT=$v python scripts/msm_bench.py 22 3 2>&1 | grep -E "ell=|accumulate|checksum" | head -3; done > gpurun_out/r2b_variants.log 2>&1; cat gpurun_out/r2b_variants.log
