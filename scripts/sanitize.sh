#!/bin/bash
# compute-sanitizer passes over small end-to-end proofs (run on a GPU box: gpurun -- bash scripts/sanitize.sh)
set -u
cd "$(dirname "$0")/.."
for tool in memcheck racecheck initcheck synccheck; do
  echo "== $tool"
  sel="point_add_flow"
  [ "$tool" = memcheck ] && sel="point_add_flow or device_builder or flow_m7 or side_stream or spmv_and_transpose"
  compute-sanitizer --tool "$tool" --error-exitcode 99 --print-limit 5 python -m pytest tests/test_gpu_prove.py tests/test_gpu_kernels.py -x -q -k "$sel" 2>&1 | tail -6
done
