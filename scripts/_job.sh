cd $GRAFT_REPO_ROOT
ncu --metrics gpu__time_duration.sum -c 1 python -c "
import os,torch
torch.zeros(1).cuda()
print('ENV', {k:v for k,v in os.environ.items() if 'INJECT' in k or 'NSIGHT' in k or k.startswith('NV') or 'PROFIL' in k or 'CUPTI' in k})
" 2>&1 | grep ENV
( time ncu --metrics gpu__time_duration.sum --csv --log-file gpurun_out/r2e_smoke_ncu.csv python -c "import __graft_entry__ as g; g.smoke()" ) 2>&1 | tail -6
grep -c "k_" gpurun_out/r2e_smoke_ncu.csv
( time python -c "import __graft_entry__ as g; g.smoke()" ) 2>&1 | tail -5
