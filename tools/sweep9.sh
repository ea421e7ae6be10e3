#!/bin/bash
run() { python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-env --groups ${G:-4} 2>/dev/null | python -c "
import json,sys,os; d=json.loads(sys.stdin.read()); print(os.environ.get('TAG',''),'ms/step %.2f sims/s %.3e'%(d['ms_per_step'],d['value']), {k:round(v['us_per_launch'],1) for k,v in d['kernels'].items() if v['us_per_launch'] and k in ('net_recurrent','backup_select')})"; }
for c in 50 25; do for tc in -1 100; do for g in 1 4 6; do
  TAG="tree_carveout=$c tc_carveout=$tc g=$g" HMZ_TREE_CARVEOUT=$c HMZ_TC_CARVEOUT=$tc G=$g run
done; done; done
