#!/bin/bash
# One gpurun call: plain bench, ncu launch list of one timed step, ncu --set full of the hot kernels.
# Outputs land in gpurun_out/ (scratch); summaries are copied to profiles/ afterwards.
# HMZ groups = 1 so that launches are serial (ncu serialises them anyway).
set -x
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-env --groups 1"
$CMD > gpurun_out/plain.json 2> gpurun_out/plain.log || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -s 620 -c 230 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_list.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:net_tc -s 111 -c 2 -o gpurun_out/prof_net $CMD > gpurun_out/ncu_net.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:search_backup_select -s 110 -c 2 -o gpurun_out/prof_tree $CMD > gpurun_out/ncu_tree.log 2>&1
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_default.json 2> gpurun_out/bench_default.log
ls -la gpurun_out/
