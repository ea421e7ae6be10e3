#!/bin/bash
run() { python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-env --groups ${G:-4} 2>/dev/null | python -c "
import json,sys,os; d=json.loads(sys.stdin.read()); print(os.environ.get('TAG',''),'ms/step %.2f sims/s %.3e'%(d['ms_per_step'],d['value']), {k:round(v['us_per_launch'],1) for k,v in d['kernels'].items() if v['us_per_launch'] and k in ('select','net_recurrent','backup_select')})"; }
TAG="g=4" G=4 run
TAG="g=1" G=1 run
