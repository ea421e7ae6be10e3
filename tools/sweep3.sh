#!/bin/bash
# A/B of the tree kernel's occupancy target: rebuilds libhmz.so on the GPU box per setting.
run() { python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-env --groups ${G:-4} 2>/dev/null | python -c "
import json,sys,os; d=json.loads(sys.stdin.read()); print(os.environ.get('TAG',''),'ms/step %.2f sims/s %.3e'%(d['ms_per_step'],d['value']), {k:round(v['us_per_launch'],1) for k,v in d['kernels'].items() if v['us_per_launch'] and k in ('select','net_recurrent','backup_select')})"; }
for mb in ${MBS:-4 5 6 8}; do
  HMZ_NVCC_EXTRA="-DHMZ_TREE_MIN_BLOCKS=$mb -DHMZ_PREFETCH_SECTORS=0 $EXTRA" python muzero-hanoi_b200/build.py --force > /dev/null 2>&1 || echo build failed
  grep -A2 "search_backup_selectILb0" muzero-hanoi_b200/build/hmz_tree.o.log | grep -i "registers\|spill" | tr '\n' ' '; echo
  for g in ${GS:-1 2 4}; do TAG="min_blocks=$mb g=$g" G=$g run; done
done
HMZ_NVCC_EXTRA="" python muzero-hanoi_b200/build.py --force > /dev/null 2>&1
