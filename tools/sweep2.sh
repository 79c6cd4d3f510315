#!/bin/bash
out=gpurun_out/sweep2.jsonl; : > $out
for tile in 256 512 1024 2048 4096; do
  CY2P_ROWS_PER_CTA=32 CY2P_TILE=$tile python tools/k1_sweep.py --B 4096 --iters 40 --tag tile >> $out 2>>gpurun_out/sweep2.err
done
CY2P_ROWS_PER_CTA=32 CY2P_TILE=1024 python tools/k1_sweep.py --B 4096 --iters 40 --eps 0 --tag noeps >> $out 2>>gpurun_out/sweep2.err
CY2P_ROWS_PER_CTA=16 CY2P_TILE=1024 python tools/k1_sweep.py --B 4096 --iters 40 --tag rpc16 >> $out 2>>gpurun_out/sweep2.err
CY2P_ROWS_PER_CTA=64 CY2P_TILE=1024 python tools/k1_sweep.py --B 4096 --iters 40 --tag rpc64 >> $out 2>>gpurun_out/sweep2.err
CY2P_ROWS_PER_CTA=32 CY2P_TILE=1024 python tools/k1_sweep.py --B 4096 --iters 40 --order rcm --tag rcm >> $out 2>>gpurun_out/sweep2.err
CY2P_ROWS_PER_CTA=32 CY2P_TILE=512 python tools/k1_sweep.py --B 4096 --iters 40 --f64 1 --tag f64 >> $out 2>>gpurun_out/sweep2.err
cat $out
