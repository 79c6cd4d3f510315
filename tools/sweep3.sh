#!/bin/bash
out=gpurun_out/sweep3.jsonl; : > $out
for ord in none rcm; do
  python tools/k1_sweep.py --n 250000 --B 1024 --iters 20 --order $ord --tag n250k >> $out 2>>gpurun_out/sweep3.err
done
python tools/k1_sweep.py --n 50000 --B 2048 --iters 40 --order rcm --tag n50k_rcm >> $out 2>>gpurun_out/sweep3.err
python tools/k1_sweep.py --n 50000 --B 2048 --iters 40 --order none --tag n50k_none >> $out 2>>gpurun_out/sweep3.err
cat $out
