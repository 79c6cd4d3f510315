#!/bin/bash
out=gpurun_out/config5_fused.jsonl; : > $out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29544"
timeout 400 $TR tools/bench_config5.py --B 1 --iters 200 --fused 1 >> $out 2>gpurun_out/config5_fused.err
timeout 400 $TR tools/bench_config5.py --B 256 --iters 40 --fused 1 >> $out 2>>gpurun_out/config5_fused.err
timeout 400 $TR tools/bench_config5.py --B 32 --iters 100 --fused 1 >> $out 2>>gpurun_out/config5_fused.err
timeout 400 $TR tools/bench_config5.py --B 32 --iters 100 --fused 0 >> $out 2>>gpurun_out/config5_fused.err
python tools/bench_config5.py --B 32 --iters 100 >> $out 2>>gpurun_out/config5_fused.err
cat $out
