#!/bin/bash
run() { python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-env --groups ${G:-4} 2>/dev/null | python -c "
import json,sys,os; d=json.loads(sys.stdin.read()); print(os.environ.get('TAG',''),'ms/step %.2f sims/s %.3e'%(d['ms_per_step'],d['value']), {k:round(v['us_per_launch'],1) for k,v in d['kernels'].items() if v['us_per_launch'] and k in ('net_recurrent','backup_select')})"; }
for pdl in 0 1 2 3; do for g in 2 3 4 6; do TAG="HMZ_PDL=$pdl g=$g" HMZ_PDL=$pdl G=$g run; done; done
