#!/bin/bash
# HMZ_TREE_MIN_BLOCKS 5 vs 7 at smaller batches (default grouping)
run() { python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-env --games $N 2>/dev/null | python -c "
import json,sys,os; d=json.loads(sys.stdin.read()); k=d['kernels']; print(os.environ.get('TAG',''),'ms/step %.3f sims/s %.3e'%(d['ms_per_step'],d['value']), 'net %.1f us tree %.1f us'%(k['net_recurrent']['us_per_launch'],k['backup_select']['us_per_launch']))"; }
for cfg in "-DHMZ_TREE_MIN_BLOCKS=5" "-DHMZ_TREE_MIN_BLOCKS=7"; do
  HMZ_NVCC_EXTRA="$cfg" python muzero-hanoi_b200/build.py --force > /dev/null 2>&1 || echo build failed
  for n in 2048 8192 16384 32768 131072; do TAG="[$cfg] games=$n" N=$n run; done
done
HMZ_NVCC_EXTRA="" python muzero-hanoi_b200/build.py --force > /dev/null 2>&1
