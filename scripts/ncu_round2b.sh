#!/bin/bash
# Round-2 (second half) profiler evidence; run on a GPU box, outputs in gpurun_out/ (summaries are copied to profiles/ afterwards).
# Pre-launched round kernels wait for a host post while ncu serialises kernels and blocks the launching thread: every ncu pass
# therefore runs with VPIN_PRELAUNCH_Q=0 (challenges as kernel parameters - the same kernels, the same work).
mkdir -p gpurun_out
export VPIN_PRELAUNCH_Q=0
# 1. launch list of exactly ONE warm bench step (CNN A, both instances, encode on its second context, derefs row half on the side stream)
timeout 400 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2b_launches_step.csv python scripts/profile_step.py A > gpurun_out/r2b_launches_step.log 2>&1
# 2. ncu --set full of the finishing kernels of a commitment (2^22 uniform scalars, 2048 x 2048): quad Horner pass and segment sum
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"k_msm_horner_quad|k_msm_segsum|k_recode" --launch-skip 3 -c 3 -o gpurun_out/r2b_msm_finish python scripts/msm_bench.py 22 1 > gpurun_out/r2b_ncu_finish.log 2>&1
ncu -i gpurun_out/r2b_msm_finish.ncu-rep --page raw --csv > gpurun_out/r2b_ncu_msm_finish_raw.csv 2>/dev/null
ls -la gpurun_out/*.ncu-rep
