#!/bin/bash
# config 5 on 1 and 2 GPUs (run under: gpurun --gpus 2)
out=gpurun_out/config5.jsonl; : > $out
python tools/bench_config5.py --B 1 --iters 200 >> $out 2>gpurun_out/config5.err
python tools/bench_config5.py --B 256 --iters 40 >> $out 2>>gpurun_out/config5.err
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533"
$TR tools/bench_config5.py --B 1 --iters 200 >> $out 2>>gpurun_out/config5.err
$TR tools/bench_config5.py --B 256 --iters 40 >> $out 2>>gpurun_out/config5.err
cat $out
