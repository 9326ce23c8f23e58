#!/bin/bash
# round-2 tuning sweep: Lanczos warp-kernel variants, MS-SSIM pipelined two-level step
out=gpurun_out/${1:-r2_sweep}.txt; : > $out
run() { echo "== $*" | tee -a $out; env "$@" 2>&1 | grep -E '"op"' | sed 's/"items_per_s.*frac_of_measured_hbm"/"frac"/' | tee -a $out; }
V=fennec_b200/_variants
run python tools/bench_ops.py lanczos
run FB_LZ_OLD=1 python tools/bench_ops.py lanczos
for v in m24 m16 s2 s4 r8 r32; do run FB_LIB_PATH=$V/libfennec_lz_$v.so python tools/bench_ops.py lanczos; done
run python tools/bench_ops.py msssim
run FB_LIB_PATH=$V/libfennec_f2pipe.so FB_F2_MINB=2 python tools/bench_ops.py msssim
run FB_F2_MINB=2 python tools/bench_ops.py msssim
