set -x
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2q_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2q_pytest.log
tail -5 gpurun_out/r2q_pytest.log
rm -f gpurun_out/r2q_time_step.log
for q in 0 16384 131072; do
  echo "== VPIN_PRELAUNCH_Q=$q" >> gpurun_out/r2q_time_step.log
  VPIN_PRELAUNCH_Q=$q timeout 300 python scripts/time_step.py A 2>&1 | grep -v "^Exception\|^Traceback\|File\|TypeError" >> gpurun_out/r2q_time_step.log
done
cat gpurun_out/r2q_time_step.log
