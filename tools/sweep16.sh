#!/bin/bash
# Tree-kernel occupancy target with the path-element backup: 7 blocks of 128 threads per SM (72 registers) would hold
# all 1,024 blocks of a 65,536-search launch in ONE wave (148 x 7 = 1,036 slots)
run() { python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-env --groups ${G:-4} 2>/dev/null | python -c "
import json,sys,os; d=json.loads(sys.stdin.read()); k=d['kernels']; print(os.environ.get('TAG',''),'ms/step %.3f sims/s %.3e'%(d['ms_per_step'],d['value']), 'net %.1f us tree %.1f us'%(k['net_recurrent']['us_per_launch'],k['backup_select']['us_per_launch']))"; }
for cfg in "-DHMZ_TREE_MIN_BLOCKS=5" "-DHMZ_TREE_MIN_BLOCKS=6" "-DHMZ_TREE_MIN_BLOCKS=7" "-DHMZ_TREE_MIN_BLOCKS=8"; do
  HMZ_NVCC_EXTRA="$cfg" python muzero-hanoi_b200/build.py --force > /dev/null 2>&1 || echo build failed
  for g in 1 2 4; do TAG="[$cfg] g=$g" G=$g run; done
done
HMZ_NVCC_EXTRA="" python muzero-hanoi_b200/build.py --force > /dev/null 2>&1
