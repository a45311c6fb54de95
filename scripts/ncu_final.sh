#!/bin/bash
# Round-2 profiler evidence (run on a GPU box; reports land in gpurun_out/, summaries are copied to profiles/ afterwards).
mkdir -p gpurun_out
# 1. ncu --set full of the dominant kernel: k_msm_accumulate on 2^22 uniform full-width scalars
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_msm_accumulate$ -c 1 -o gpurun_out/r2_msm_acc python scripts/msm_bench.py 22 1 > gpurun_out/r2_ncu_msm.log 2>&1
# 2. launch list of exactly ONE warm bench step (CNN A, both instances)
timeout 400 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_step.csv python scripts/profile_step.py A > gpurun_out/r2_launches_step.log 2>&1
# 3. DRAM bytes and duration of every k_msm_accumulate launch of that step (roofline.traffic of bench.py)
timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed --clock-control none -k regex:k_msm_accumulate$ --csv --log-file gpurun_out/r2_msm_step_traffic.csv python scripts/profile_step.py A > gpurun_out/r2_msm_step_traffic.log 2>&1
# 4. ncu --set full of the two largest batched sumcheck rounds of a CNN-E proof (launches 120, 121 of the kernel: the leaf layer of the ops proof)
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_round_cubic_batched\< --launch-skip 120 -c 2 -o gpurun_out/r2_round_batched_E python scripts/prove_shape.py E 1 > gpurun_out/r2_ncu_round.log 2>&1
ls -la gpurun_out/*.ncu-rep
