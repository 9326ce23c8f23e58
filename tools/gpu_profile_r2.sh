#!/bin/bash
# usage: tools/gpu_profile_r2.sh <tag> <op> [<op> ...]  — ncu --set full of the op's kernels with opaque alpha (photo case); source + raw CSV pages
tag=$1; shift
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on"
for op in "$@"; do
  drv=$op
  case $op in
    lanczos) args="--pairs 2 --iters 1 --w 7680 --h 4320 --opaque"; rx='resize|lanczos'; cnt=2;;
    lanczos_translucent) args="--pairs 2 --iters 1 --w 7680 --h 4320"; rx='resize|lanczos'; cnt=2; drv=lanczos;;
    msssim)  args="--pairs 2 --iters 1 --w 7680 --h 4320"; rx='box|ssim'; cnt=14;;
    ssim)    args="--pairs 32 --iters 2"; rx='ssim_strip'; cnt=1;;
    blur)    args="--pairs 4 --iters 1 --opaque"; rx="blur"; cnt=2;;
    sharpen) args="--pairs 4 --iters 1 --opaque"; rx="fx_tile"; cnt=1;;
    *)       args="--pairs 4 --iters 1 --opaque"; rx="blur|sharpen|fx_tile|adaptive|box|ycbcr|analyze|orient|palette"; cnt=4;;
  esac
  timeout 300 $NCU -k regex:$rx -c $cnt -f -o gpurun_out/${tag}_$op python tools/profile_driver.py $drv $args > gpurun_out/${tag}_$op.log 2>&1
  ncu -i gpurun_out/${tag}_$op.ncu-rep --page raw --csv > gpurun_out/${tag}_$op.raw.csv 2>/dev/null
  ncu -i gpurun_out/${tag}_$op.ncu-rep --page source --csv --print-source sass > gpurun_out/${tag}_$op.source.csv 2>/dev/null
done
ls -la gpurun_out | tail -20
