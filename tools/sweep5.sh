#!/bin/bash
out=gpurun_out/sweep5.jsonl; : > $out
for aff in 0 1; do
  CY2P_SM_AFFINE=$aff python tools/k1_sweep.py --n 50000 --B 4096 --iters 40 --tag aff$aff >> $out 2>>gpurun_out/sweep5.err
  CY2P_SM_AFFINE=$aff python tools/k1_sweep.py --n 250000 --B 1024 --iters 20 --tag aff$aff >> $out 2>>gpurun_out/sweep5.err
done
CY2P_SM_AFFINE=1 CY2P_ROWS_PER_CTA=16 python tools/k1_sweep.py --n 50000 --B 4096 --iters 40 --tag aff1_rpc16 >> $out 2>>gpurun_out/sweep5.err
CY2P_SM_AFFINE=1 CY2P_ROWS_PER_CTA=64 python tools/k1_sweep.py --n 50000 --B 4096 --iters 40 --tag aff1_rpc64 >> $out 2>>gpurun_out/sweep5.err
cat $out
