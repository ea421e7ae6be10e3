#!/bin/bash
# Where the kernels of the search loop signal their programmatic dependents (HMZ_PDL_NET_AT / HMZ_PDL_TREE_AT), same box
run() { python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-env --groups ${G:-4} 2>/dev/null | python -c "
import json,sys,os; d=json.loads(sys.stdin.read()); print(os.environ.get('TAG',''),'ms/step %.3f sims/s %.3e'%(d['ms_per_step'],d['value']))"; }
for rep in 1 2; do
for cfg in "7 -1 0" "7 3 0" "7 2 0" "7 -1 2" "7 3 2" "7 2 2" "7 3 1" "7 4 3" "3 -1 2" "7 3 3"; do
  set -- $cfg
  for g in 1 4; do export HMZ_PDL=$1 HMZ_PDL_NET_AT=$2 HMZ_PDL_TREE_AT=$3; TAG="pdl=$1 net_at=$2 tree_at=$3 g=$g" G=$g run; done
done
done
