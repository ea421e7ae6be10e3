#!/bin/bash
# Ablation timing of the fused tree kernel (results are WRONG in ablated builds; timing only).
run() { python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-env --groups 1 2>/dev/null | python -c "
import json,sys,os; d=json.loads(sys.stdin.read()); print(os.environ.get('TAG',''),'ms/step %.2f'%(d['ms_per_step']), {k:round(v['us_per_launch'],1) for k,v in d['kernels'].items() if v['us_per_launch'] and k in ('net_recurrent','backup_select')})"; }
for ab in 0 1 2 4 3 6 5; do
  HMZ_NVCC_EXTRA="-DHMZ_TREE_MIN_BLOCKS=5 -DHMZ_PREFETCH_SECTORS=0 -DHMZ_ABLATE=$ab" python muzero-hanoi_b200/build.py --force > /dev/null 2>&1 || echo build failed
  TAG="ablate=$ab (1: no select math, 2: no backup, 4: no select)" run
done
HMZ_NVCC_EXTRA="" python muzero-hanoi_b200/build.py --force > /dev/null 2>&1
