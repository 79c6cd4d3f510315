#!/bin/bash
# ncu source-level capture (dense warp sampling) of the node-level kernels of cy2p_fhmm_loss_grad at config-3 shapes.
# usage (under gpurun): tools/profile_fhmm_nodes.sh <tag>  -> gpurun_out/r02/fhmm_nodes_<tag>*
set -x
TAG=${1:-a}
OUT=gpurun_out/r02
mkdir -p $OUT
CMD="python tools/bench_fhmm_mma.py --only restricted_mma_serial --reps 3"
ncu --set full --warp-sampling-interval 1 --clock-control none --import-source on \
    -k regex:"fhmm_node_pass_c|fhmm_node_pass_b|fhmm_node_pass_a|spmm_gather|csr_times_small" -s 10 -c 5 -f \
    -o $OUT/fhmm_nodes_$TAG $CMD > $OUT/fhmm_nodes_ncu_$TAG.log 2>&1
ncu -i $OUT/fhmm_nodes_$TAG.ncu-rep --page raw --csv > $OUT/fhmm_nodes_${TAG}_raw.csv 2>/dev/null
ncu -i $OUT/fhmm_nodes_$TAG.ncu-rep --page source --csv > $OUT/fhmm_nodes_${TAG}_source.csv 2>/dev/null
python tools/ncu_summary.py $OUT/fhmm_nodes_${TAG}_raw.csv > $OUT/fhmm_nodes_${TAG}_summary.json
tail -3 $OUT/fhmm_nodes_ncu_$TAG.log
