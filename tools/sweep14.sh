#!/bin/bash
# Tree / MLP kernel duration vs searches per launch (one group): is the tree kernel's 28 us at 65,536 searches a wave
# effect (1,024 blocks on 148 x 5 = 740 slots) or a throughput limit?
for n in 8192 16384 32768 47360 49152 57344 65536; do
  python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-env --groups 1 --games $n 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); k=d['kernels']; print('games $n', 'ms/step %.3f sims/s %.3e'%(d['ms_per_step'],d['value']), 'net %.1f us tree %.1f us'%(k['net_recurrent']['us_per_launch'],k['backup_select']['us_per_launch']))"
done
