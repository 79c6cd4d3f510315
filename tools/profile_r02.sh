#!/bin/bash
# ncu evidence for profiles/ (run under gpurun, one GPU). Each ncu run follows a plain run of the same command.
# usage: tools/profile_r02.sh <tag>      -> gpurun_out/r02/{launches,ncu_k1}_<tag>*
set -x
TAG=${1:-r02b}
OUT=gpurun_out/r02
mkdir -p $OUT
CMD="python bench.py --workload profile --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-extras"
$CMD > $OUT/plain_profile_$TAG.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_$TAG.csv $CMD > $OUT/ncu_launches_$TAG.log 2>&1
# the tile -> tile update (template flags IN_F32 = 0, OUT_F32 = 0): skip the plan kernels and the first update
ncu --set full --clock-control none --import-source on -k regex:spmm_gather_x -s 3 -c 2 -f -o $OUT/prof_k1_$TAG $CMD > $OUT/ncu_k1_$TAG.log 2>&1
ncu -i $OUT/prof_k1_$TAG.ncu-rep --page raw --csv > $OUT/ncu_k1_${TAG}_raw.csv 2>/dev/null
python tools/ncu_summary.py $OUT/ncu_k1_${TAG}_raw.csv > $OUT/ncu_k1_${TAG}_summary.json
ls -la $OUT
