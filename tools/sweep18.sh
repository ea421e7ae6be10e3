#!/bin/bash
# Fat tree blocks: does packing the tree kernel onto few SMs (one 640-thread block = a whole register file) leave
# whole SMs to the tensor-core kernels of the other groups?
run() { python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-env --groups ${G:-4} 2>/dev/null | python -c "
import json,sys,os; d=json.loads(sys.stdin.read()); k=d['kernels']; print(os.environ.get('TAG',''),'ms/step %.3f sims/s %.3e'%(d['ms_per_step'],d['value']), 'net %.1f us tree %.1f us'%(k['net_recurrent']['us_per_launch'],k['backup_select']['us_per_launch']))"; }
for cfg in "-DHMZ_TREE_THREADS=128 -DHMZ_TREE_MIN_BLOCKS=5" "-DHMZ_TREE_THREADS=320 -DHMZ_TREE_MIN_BLOCKS=2" "-DHMZ_TREE_THREADS=640 -DHMZ_TREE_MIN_BLOCKS=1" "-DHMZ_TREE_THREADS=512 -DHMZ_TREE_MIN_BLOCKS=1" "-DHMZ_TREE_THREADS=1024 -DHMZ_TREE_MIN_BLOCKS=1"; do
  HMZ_NVCC_EXTRA="$cfg" python muzero-hanoi_b200/build.py --force > /dev/null 2>&1 || echo build failed
  for g in 1 4 8; do TAG="[$cfg] g=$g" G=$g run; done
done
HMZ_NVCC_EXTRA="" python muzero-hanoi_b200/build.py --force > /dev/null 2>&1
