#!/bin/bash
out=gpurun_out/sweep7b.jsonl; : > $out
for cfg in "32 64" "32 128" "64 128" "128 128" "128 64" "32 96" "64 64"; do
  set -- $cfg
  env CY2P_GROUP=$1 CY2P_ROWS_PER_CTA=$2 python tools/k1_sweep.py --n 50000 --B 4096 --iters 40 --tag g$1_r$2 >> $out 2>>gpurun_out/sweep7.err
done
env CY2P_GROUP=32 CY2P_ROWS_PER_CTA=64 python tools/k1_sweep.py --n 250000 --B 1024 --iters 20 --tag g32_r64_250k >> $out 2>>gpurun_out/sweep7.err
env CY2P_GROUP=64 CY2P_ROWS_PER_CTA=128 python tools/k1_sweep.py --n 250000 --B 1024 --iters 20 --tag g64_r128_250k >> $out 2>>gpurun_out/sweep7.err
cat $out
