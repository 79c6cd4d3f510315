#!/bin/bash
# warps per CTA of the gather kernel (library variants built with kWarps = 4 / 16 next to the default 8)
out=gpurun_out/sweep8.jsonl; : > $out
run() { tag=$1; shift; env "$@" python tools/k1_sweep.py --n 50000 --B 4096 --iters 40 --tag $tag >> $out 2>>gpurun_out/sweep8.err; }
run kw8_r64
cp cy2path_b200/libcy2path_b200.so /tmp/lib_default.so
cp cy2path_b200/build/lib_kw16.so cy2path_b200/libcy2path_b200.so
run kw16_r64
run kw16_r128 CY2P_ROWS_PER_CTA=128
run kw16_r32 CY2P_ROWS_PER_CTA=32
cp cy2path_b200/build/lib_kw4.so cy2path_b200/libcy2path_b200.so
run kw4_r64
run kw4_r32 CY2P_ROWS_PER_CTA=32
run kw4_r16 CY2P_ROWS_PER_CTA=16
cp /tmp/lib_default.so cy2path_b200/libcy2path_b200.so
cat $out
