#!/bin/bash
# K1 experiments: every row-walk / block-shape variant of ssim_strip_kernel on 64 4K pairs (parity on the golden cases first).
out=gpurun_out/${1:-k1_variants}.txt
mkdir -p $(dirname $out)
: > $out
run() { echo "== $*" | tee -a $out; env "$@" PAIRS=64 python tools/quick_ssim.py 2>&1 | tail -2 | tee -a $out; }
if [ -n "$2" ]; then shift; for v in "$@"; do run $v; done; exit 0; fi
run FB_SSIM_MODE=0 FB_SSIM_WPB=4
run FB_SSIM_MODE=0 FB_SSIM_WPB=1
run FB_SSIM_MODE=2 FB_SSIM_WPB=4
run FB_SSIM_MODE=2 FB_SSIM_WPB=1
run FB_SSIM_MODE=3
run FB_SSIM_MODE=4
run FB_SSIM_MODE=2 FB_SSIM_REGCAP=184
run FB_SSIM_MODE=3 FB_SSIM_REGCAP=184
run FB_SSIM_MODE=4 FB_SSIM_REGCAP=184
run FB_SSIM_MODE=4 FB_SSIM_REGCAP=168
