#!/bin/bash
# round-2 tuning sweep over the variant builds of tools/build_variant.sh
out=gpurun_out/${1:-r2_sweep}.txt; : > $out
run() { echo "== $*" | tee -a $out; env "$@" 2>&1 | grep -E '"op"' | grep -vE "boxDownsample|Adaptive" | sed 's/"items_per_s.*frac_of_measured_hbm"/"frac"/' | tee -a $out; }
V=fennec_b200/_variants
run python tools/bench_ops.py msssim
for v in f2pf1 f2pf2 f2pf3; do run FB_LIB_PATH=$V/libfennec_$v.so python tools/bench_ops.py msssim; done
run python tools/bench_ops.py blur
for v in bl_v120 bl_v360 bl_v540 bl_h16 bl_h4; do run FB_LIB_PATH=$V/libfennec_$v.so python tools/bench_ops.py blur; done
run python tools/bench_ops.py lanczos
for v in lz_r24 lz_r32 lz_v6 lz_v12 lz_m24; do run FB_LIB_PATH=$V/libfennec_$v.so python tools/bench_ops.py lanczos; done
