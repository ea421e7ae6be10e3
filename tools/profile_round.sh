#!/bin/bash
# One gpurun call: plain bench, ncu launch list of one timed step, ncu --set full of the three
# hot kernels.  Outputs land in gpurun_out/ (scratch); summaries are copied to profiles/ by hand.
set -x
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-env"
$CMD > gpurun_out/plain.json 2> gpurun_out/plain.log || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -s 925 -c 330 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_list.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:net_recurrent_tc -s 310 -c 2 -o gpurun_out/prof_net $CMD > gpurun_out/ncu_net.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:search_select -s 310 -c 2 -o gpurun_out/prof_select $CMD > gpurun_out/ncu_select.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:search_expand_backup -s 310 -c 2 -o gpurun_out/prof_backup $CMD > gpurun_out/ncu_backup.log 2>&1
ls -la gpurun_out/
