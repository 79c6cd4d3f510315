#!/bin/bash
set -x
CMD="python tools/bench_fhmm.py --epochs 3"
$CMD > gpurun_out/plain_fhmm.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 300 --csv --log-file gpurun_out/launches_fhmm_r01b.csv $CMD > gpurun_out/ncu_fhmm_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:fhmm_data_pass -s 3 -c 1 -f -o gpurun_out/prof_fhmm_data_r01b $CMD > gpurun_out/ncu_fhmm_data.log 2>&1
