#!/bin/bash
# usage: tools/gpu_profile_some.sh <op> [<op> ...]   (ops of tools/profile_driver.py; 8K for lanczos/msssim)
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on"
for op in "$@"; do
  case $op in
    lanczos) args="--pairs 2 --iters 1 --w 7680 --h 4320"; rx='resize|lanczos'; cnt=4;;
    msssim)  args="--pairs 2 --iters 1 --w 7680 --h 4320"; rx='box|ssim'; cnt=14;;
    ssim)    args="--pairs 32 --iters 2"; rx='ssim_strip'; cnt=1;;
    *)       args="--pairs 4 --iters 1"; rx="blur|sharpen|fx_tile|adaptive|box|ycbcr|analyze|orient|palette"; cnt=4;;
  esac
  timeout 240 $NCU -k regex:$rx -c $cnt -f -o gpurun_out/$op python tools/profile_driver.py $op $args > gpurun_out/$op.log 2>&1
done
ls -la gpurun_out | tail -20
