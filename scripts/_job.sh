cd $GRAFT_REPO_ROOT
VARIANTS="0 20 0 20" bash scripts/run_msm_variants.sh > /dev/null 2>&1
cp gpurun_out/r2_msm_variants.log gpurun_out/r2u_msm_pf.log; cat gpurun_out/r2u_msm_pf.log
