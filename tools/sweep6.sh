#!/bin/bash
# staged gather vs plain gather, grouped vs plain Cuthill-McKee order
out=gpurun_out/sweep6.jsonl; : > $out
run() { tag=$1; shift; env "$@" python tools/k1_sweep.py --n 50000 --B 4096 --iters 40 --tag $tag >> $out 2>>gpurun_out/sweep6.err; }
run grp_plain CY2P_GROUP=32 CY2P_STAGED=0
run staged_u272_s8 CY2P_STAGED=1
run staged_u272_s16 CY2P_STAGED=1 CY2P_STAGE_SLABS=16
run staged_u272_s32 CY2P_STAGED=1 CY2P_STAGE_SLABS=32
run staged_u200_s8 CY2P_STAGED=1 CY2P_STAGE_UCAP=200
run staged_u384_s8 CY2P_STAGED=1 CY2P_STAGE_UCAP=384 CY2P_STAGE_ROWS=48 CY2P_GROUP=48
run staged_u160_s8 CY2P_STAGED=1 CY2P_STAGE_UCAP=160
cat $out
