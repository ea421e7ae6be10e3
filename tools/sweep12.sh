#!/bin/bash
run() { python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-env --groups ${G:-4} 2>/dev/null | python -c "
import json,sys,os; d=json.loads(sys.stdin.read()); print(os.environ.get('TAG',''),'ms/step %.2f sims/s %.3e'%(d['ms_per_step'],d['value']), {k:round(v['us_per_launch'],1) for k,v in d['kernels'].items() if v['us_per_launch'] and k in ('net_recurrent','backup_select')})"; }
for cfg in "" "-DHMZ_NO_LATENT_PREFETCH"; do
  HMZ_NVCC_EXTRA="$cfg" python muzero-hanoi_b200/build.py --force > /dev/null 2>&1 || echo build failed
  for g in 1 4; do TAG="[$cfg] g=$g" G=$g run; done
done
HMZ_NVCC_EXTRA="" python muzero-hanoi_b200/build.py --force > /dev/null 2>&1
