# measured alternatives of the MSM hot loop (VPIN_MSM_VARIANT, kernels_msm.cu); run on a GPU box
mkdir -p gpurun_out
for v in ${VARIANTS:-0 1 2 12}; do echo "== variant $v"; VPIN_MSM_VARIANT=$v python scripts/msm_bench.py 22 3 2>&1 | grep -E "ell=|accumulate|checksum" | head -3; done > gpurun_out/r2_msm_variants.log 2>&1; cat gpurun_out/r2_msm_variants.log
