#!/bin/bash
# ncu --set full of the kernels added late in round 2: x_eps_reduce, sample_chains_prefetch (config-2 profile workload)
# and spmm_rows_narrow (config-3 loss+grad). usage (under gpurun): tools/profile_r02_new_kernels.sh
set -x
OUT=gpurun_out/r02
mkdir -p $OUT
CMD="python bench.py --workload profile --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-extras"
ncu --set full --clock-control none -k regex:"x_eps_reduce|sample_chains_prefetch" -s 4 -c 3 -f -o $OUT/new_k1 $CMD > $OUT/new_k1.log 2>&1
ncu -i $OUT/new_k1.ncu-rep --page raw --csv > $OUT/new_k1_raw.csv 2>/dev/null
python tools/ncu_summary.py $OUT/new_k1_raw.csv > $OUT/new_k1_summary.json
CMD2="python tools/bench_fhmm_mma.py --only restricted_mma_serial --reps 3"
ncu --set full --clock-control none -k regex:"spmm_rows_narrow|csr_times_small" -s 2 -c 2 -f -o $OUT/new_k3 $CMD2 > $OUT/new_k3.log 2>&1
ncu -i $OUT/new_k3.ncu-rep --page raw --csv > $OUT/new_k3_raw.csv 2>/dev/null
python tools/ncu_summary.py $OUT/new_k3_raw.csv > $OUT/new_k3_summary.json
