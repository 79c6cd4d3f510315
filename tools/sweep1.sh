#!/bin/bash
# K1 tuning sweep (one process per setting; knobs are read once per process)
out=gpurun_out/sweep1.jsonl; : > $out
for rpc in 8 16 32 64; do for tile in 128 256 512; do
  CY2P_ROWS_PER_CTA=$rpc CY2P_TILE=$tile python tools/k1_sweep.py --B 1024 --iters 100 --tag rpc_tile >> $out 2>>gpurun_out/sweep1.err
done; done
CY2P_ROWS_PER_CTA=32 CY2P_TILE=128 python tools/k1_sweep.py --B 1024 --iters 100 --eps 0 --tag noeps >> $out 2>>gpurun_out/sweep1.err
CY2P_ROWS_PER_CTA=32 CY2P_TILE=256 python tools/k1_sweep.py --B 1024 --iters 100 --eps 0 --tag noeps >> $out 2>>gpurun_out/sweep1.err
for ord in blob rcm tau; do
  CY2P_ROWS_PER_CTA=32 CY2P_TILE=128 python tools/k1_sweep.py --B 1024 --iters 100 --order $ord --tag order >> $out 2>>gpurun_out/sweep1.err
done
CY2P_ROWS_PER_CTA=64 CY2P_TILE=128 python tools/k1_sweep.py --B 1024 --iters 100 --order blob --tag order64 >> $out 2>>gpurun_out/sweep1.err
CY2P_ROWS_PER_CTA=32 CY2P_TILE=128 python tools/k1_sweep.py --B 1024 --iters 100 --f64 1 --tag f64 >> $out 2>>gpurun_out/sweep1.err
CY2P_ROWS_PER_CTA=32 CY2P_TILE=64 python tools/k1_sweep.py --B 1024 --iters 100 --f64 1 --tag f64 >> $out 2>>gpurun_out/sweep1.err
cat $out
