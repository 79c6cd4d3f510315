#!/bin/bash
# config 5: halo push vs fused all-gather vs NCCL all-gather (run under: gpurun --gpus G; G from $1, default 2)
G=${1:-2}
out=gpurun_out/config5_halo_${G}gpu.jsonl; : > $out
err=gpurun_out/config5_halo.err; : > $err
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29555"
python tools/bench_config5.py --B 256 --iters 40 2>>$err | grep '^{' >> $out   # 1 GPU reference; builds the graph cache
python tools/bench_config5.py --B 1 --iters 200 2>>$err | grep '^{' >> $out
for B in 256 1; do
  it=40; [ $B -eq 1 ] && it=200
  timeout 400 $TR tools/bench_config5.py --B $B --iters $it --halo 1 2>>$err | grep '^{' >> $out
  timeout 400 $TR tools/bench_config5.py --B $B --iters $it --fused 1 2>>$err | grep '^{' >> $out
  timeout 400 $TR tools/bench_config5.py --B $B --iters $it 2>>$err | grep '^{' >> $out
done
cat $out; tail -3 $err
