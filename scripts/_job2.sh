cd $GRAFT_REPO_ROOT
rm -f gpurun_out/r2x_early2.log
for cfg in "1 1" "1 2" "1 3"; do set -- $cfg
  echo "== VPIN_DEREFS_EARLY=$1 VPIN_SIDE_MSM_BLOCKS=$2" >> gpurun_out/r2x_early2.log
  VPIN_DEREFS_EARLY=$1 VPIN_SIDE_MSM_BLOCKS=$2 timeout 300 python scripts/time_step.py A 2>&1 | grep -A1 "^point_mult" | sed 's/eval_sparse.*commit_nondet/... commit_nondet/; s/network_alloc.*SNARK/... SNARK/' >> gpurun_out/r2x_early2.log
done
cat gpurun_out/r2x_early2.log
