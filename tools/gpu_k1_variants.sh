#!/bin/bash
# K1 experiments: every row-walk / block-shape variant of ssim_strip_kernel on 64 4K pairs (parity on the golden cases first).
out=gpurun_out/${1:-k1_variants}.txt
: > $out
for v in "FB_SSIM_MODE=0 FB_SSIM_WPB=4" "FB_SSIM_MODE=0 FB_SSIM_WPB=1" "FB_SSIM_MODE=2 FB_SSIM_WPB=4" "FB_SSIM_MODE=2 FB_SSIM_WPB=1" "FB_SSIM_MODE=1 FB_SSIM_WPB=4"; do
  echo "== $v" | tee -a $out
  env $v PAIRS=64 python tools/quick_ssim.py 2>&1 | tail -3 | tee -a $out
done
