#!/bin/bash
# end-of-round measurements (one GPU): configs 1, 4, model training, decoders
python bench.py --workload config1 > gpurun_out/bench_config1_r01_d.json 2> gpurun_out/bench_config1.err
python bench.py --workload config4 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_config4_r01_d.json 2> gpurun_out/bench_config4.err
: > gpurun_out/bench_models_r01.jsonl
for m in FHMM LSSM NSSM; do
  python tools/bench_fhmm.py --model $m >> gpurun_out/bench_models_r01.jsonl 2>> gpurun_out/bench_models.err
  python tools/bench_fhmm.py --model $m --unrestricted >> gpurun_out/bench_models_r01.jsonl 2>> gpurun_out/bench_models.err
done
tail -c 600 gpurun_out/bench_config1_r01_d.json; echo; tail -c 600 gpurun_out/bench_config4_r01_d.json; echo; cat gpurun_out/bench_models_r01.jsonl
