#!/bin/bash
# ncu --set full captures of the two dominant kernels (run on a GPU box; reports land in gpurun_out/, keep them under 64 MiB)
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_msm_accumulate$ -c 1 -o gpurun_out/v15_msm_acc python scripts/msm_bench.py 22 1 > gpurun_out/v15_ncu_msm.log 2>&1
# the largest batched rounds of a CNN-A proof: launches 40.. of the kernel are the leaf-layer rounds of the ops proof
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_round_cubic_batched -s 56 -c 8 -o gpurun_out/v15_round_batched python scripts/prove_shape.py A 1 > gpurun_out/v15_ncu_round.log 2>&1
ls -la gpurun_out/*.ncu-rep
