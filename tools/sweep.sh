#!/bin/bash
# quick A/B of scheduling switches: prints ms/step and per-kernel us for each setting
run() { python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-env --groups ${G:-4} 2>/dev/null | python -c "
import json,sys,os; d=json.loads(sys.stdin.read()); print(os.environ.get('TAG',''),'ms/step %.2f sims/s %.3e'%(d['ms_per_step'],d['value']), {k:round(v['us_per_launch'],1) for k,v in d['kernels'].items() if v['us_per_launch'] and k in ('select','net_recurrent','expand_backup')})"; }
TAG="tps=1 g=4" HMZ_TPS=1 run
TAG="tps=1 g=1" G=1 HMZ_TPS=1 run
TAG="tps=1 g=2" G=2 HMZ_TPS=1 run
TAG="tps=0 fused=1 g=4" HMZ_TPS=0 HMZ_FUSED=1 run
