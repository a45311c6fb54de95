cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2d_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2d_pytest.log; tail -3 gpurun_out/r2d_pytest.log
VPIN_MSM_SUB=2 VPIN_MSM_W=13 timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2d_pytest_sub2.log 2>&1; echo "pytest sub2 rc=$?" >> gpurun_out/r2d_pytest_sub2.log; tail -3 gpurun_out/r2d_pytest_sub2.log
VPIN_MSM_SUB=2 VPIN_MSM_W=15 timeout 900 python -m pytest tests/test_gpu_golden.py tests/test_gpu_kernels.py -m gpu -x -q > gpurun_out/r2d_pytest_sub2w15.log 2>&1; echo "pytest sub2 w15 rc=$?" >> gpurun_out/r2d_pytest_sub2w15.log; tail -3 gpurun_out/r2d_pytest_sub2w15.log
python scripts/msm_bench.py 22 3 2>&1 | grep -E "ell=|accumulate|finish|recode" 
timeout 600 python scripts/prove_shape_resident.py L5 3 > gpurun_out/r2d_resident_L5_1gpu.log 2>&1; tail -4 gpurun_out/r2d_resident_L5_1gpu.log
