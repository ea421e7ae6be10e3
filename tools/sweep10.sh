#!/bin/bash
run() { python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-env --groups ${G:-4} 2>&1 | tail -1 | python -c "
import json,sys,os; d=json.loads(sys.stdin.read()); print(os.environ.get('TAG',''),'ms/step %.2f sims/s %.3e e2e %.3e launches %d'%(d['ms_per_step'],d['value'],d['e2e']['value'],d['gpu_launches']))"; }
for gr in 0 1; do for g in 1 4; do TAG="HMZ_GRAPH=$gr g=$g" HMZ_GRAPH=$gr G=$g run; done; done
HMZ_GRAPH=1 timeout 600 python -m pytest tests/test_mcts_gpu.py tests/test_search_gpu.py -x -q 2>&1 | tail -3
