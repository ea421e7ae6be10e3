#!/bin/bash
# Root level (float64 priors) peeled out of the walk loop vs one level body
run() { python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-env --groups ${G:-4} 2>/dev/null | python -c "
import json,sys,os; d=json.loads(sys.stdin.read()); k=d['kernels']; print(os.environ.get('TAG',''),'ms/step %.3f sims/s %.3e'%(d['ms_per_step'],d['value']), 'net %.1f us tree %.1f us'%(k['net_recurrent']['us_per_launch'],k['backup_select']['us_per_launch']))"; }
for cfg in "-DHMZ_NO_ROOT_PEEL" "" "-DHMZ_NO_ROOT_PEEL" ""; do
  HMZ_NVCC_EXTRA="$cfg" python muzero-hanoi_b200/build.py --force > /dev/null 2>&1 || echo build failed
  TAG="[$cfg] g=4" G=4 run
done
HMZ_NVCC_EXTRA="" python muzero-hanoi_b200/build.py --force > /dev/null 2>&1
