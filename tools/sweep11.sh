#!/bin/bash
run() { python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-env --groups ${G:-4} 2>/dev/null | python -c "
import json,sys,os; d=json.loads(sys.stdin.read()); print(os.environ.get('TAG',''),'ms/step %.2f sims/s %.3e'%(d['ms_per_step'],d['value']))"; }
for cfg in "-DHMZ_TC_MAXNREG=64 -DHMZ_TREE_MIN_BLOCKS=6" "-DHMZ_TC_MAXNREG=64 -DHMZ_TREE_MIN_BLOCKS=8"; do
  HMZ_NVCC_EXTRA="$cfg" python muzero-hanoi_b200/build.py --force > /dev/null 2>&1 || echo build failed
  for c in 100 25; do for g in 4 6 8; do TAG="[$cfg] carveout=$c g=$g" HMZ_TREE_CARVEOUT=$c G=$g run; done; done
done
HMZ_NVCC_EXTRA="" python muzero-hanoi_b200/build.py --force > /dev/null 2>&1
