#!/bin/bash
# ncu evidence for profiles/ (run under gpurun, one GPU). Each ncu run follows a plain run of the same command.
set -x
CMD="python bench.py --workload profile --steps 1 --warmup 1 --no-e2e --no-cpu-baseline"
$CMD > gpurun_out/plain_profile.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r01d.csv $CMD > gpurun_out/ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:spmm_gather -s 4 -c 2 -f -o gpurun_out/prof_gather_r01d $CMD > gpurun_out/ncu_gather.log 2>&1
ls -la gpurun_out/
