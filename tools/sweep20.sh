#!/bin/bash
# Count-row table ({1/n, 1/(n+1), n, n+1} per visit count) vs clamps + two loads + two conversions in the walk
run() { python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-env --groups ${G:-4} 2>/dev/null | python -c "
import json,sys,os; d=json.loads(sys.stdin.read()); k=d['kernels']; print(os.environ.get('TAG',''),'ms/step %.3f sims/s %.3e'%(d['ms_per_step'],d['value']), 'net %.1f us tree %.1f us'%(k['net_recurrent']['us_per_launch'],k['backup_select']['us_per_launch']))"; }
for rep in 1 2; do
for cfg in "-DHMZ_NO_CNT_TABLE" ""; do
  HMZ_NVCC_EXTRA="$cfg" python muzero-hanoi_b200/build.py --force > /dev/null 2>&1 || echo build failed
  for g in 1 4; do TAG="[$cfg] g=$g" G=$g run; done
done
done
HMZ_NVCC_EXTRA="" python muzero-hanoi_b200/build.py --force > /dev/null 2>&1
