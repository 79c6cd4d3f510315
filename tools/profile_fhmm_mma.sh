#!/bin/bash
# ncu --set full capture of the three tensor-core kernels of cy2p_fhmm_loss_grad at config-3 shapes (K = 10).
# usage (under gpurun): tools/profile_fhmm_mma.sh <tag>  -> gpurun_out/r02/fhmm_mma_<tag>*
set -x
TAG=${1:-a}
OUT=gpurun_out/r02
mkdir -p $OUT
CMD="python tools/bench_fhmm_mma.py --only restricted_mma_serial --reps 3"
$CMD > $OUT/fhmm_mma_plain_$TAG.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:'fhmm_fwd_mma|fhmm_dE_mma|fhmm_dG_mma' -s 3 -c 3 -f \
    -o $OUT/fhmm_mma_$TAG $CMD > $OUT/fhmm_mma_ncu_$TAG.log 2>&1
ncu -i $OUT/fhmm_mma_$TAG.ncu-rep --page raw --csv > $OUT/fhmm_mma_${TAG}_raw.csv 2>/dev/null
python tools/ncu_summary.py $OUT/fhmm_mma_${TAG}_raw.csv > $OUT/fhmm_mma_${TAG}_summary.json
