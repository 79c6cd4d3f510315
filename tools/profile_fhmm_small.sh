#!/bin/bash
# ncu source-level capture of the single-CTA recurrence kernels of cy2p_fhmm_loss_grad at config-3 shapes.
# usage (under gpurun): tools/profile_fhmm_small.sh <tag>  -> gpurun_out/r02/fhmm_small_<tag>*
set -x
TAG=${1:-a}
OUT=gpurun_out/r02
mkdir -p $OUT
CMD="python tools/bench_fhmm_mma.py --only restricted_mma_serial --reps 3"
$CMD > $OUT/fhmm_small_plain_$TAG.log 2>&1 || exit 1
ncu --set full --warp-sampling-interval 0 --clock-control none --import-source on -k regex:"fhmm_small_forward|fhmm_backward_fused" -s 2 -c 2 -f \
    -o $OUT/fhmm_small_$TAG $CMD > $OUT/fhmm_small_ncu_$TAG.log 2>&1
ncu -i $OUT/fhmm_small_$TAG.ncu-rep --page raw --csv > $OUT/fhmm_small_${TAG}_raw.csv 2>/dev/null
ncu -i $OUT/fhmm_small_$TAG.ncu-rep --page source --csv > $OUT/fhmm_small_${TAG}_source.csv 2>/dev/null
ls -la $OUT | tail -8
