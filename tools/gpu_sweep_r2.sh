#!/bin/bash
# round-2 tuning sweep over the variant builds of tools/build_variant.sh
out=gpurun_out/${1:-r2_sweep}.txt; : > $out
run() { echo "== $*" | tee -a $out; env "$@" 2>&1 | grep -E '"op"' | grep -vE "boxDownsample|Adaptive" | sed 's/"items_per_s.*frac_of_measured_hbm"/"frac"/' | tee -a $out; }
V=fennec_b200/_variants
run python tools/bench_ops.py msssim
for v in f2p2pf0 f2p2pf1; do run FB_LIB_PATH=$V/libfennec_$v.so python tools/bench_ops.py msssim; run FB_F2_MINB=2 FB_LIB_PATH=$V/libfennec_$v.so python tools/bench_ops.py msssim; done
