#!/bin/bash
# Run on the GPU box (under gpurun): ncu --set full captures of each kernel family at BASELINE sizes + the bench launch list.
# Outputs land in gpurun_out/ (reports) — summarise here with tools/ncu_summary.py into profiles/.
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on"
run() { # name regex driver-args...
  local name=$1 rx=$2; shift 2
  timeout 240 $NCU -k regex:$rx -s ${SKIP:-0} -c ${CNT:-2} -f -o gpurun_out/$name python tools/profile_driver.py "$@" > gpurun_out/$name.log 2>&1
}
CNT=1 SKIP=1 run ssim32 ssim_strip ssim --pairs 32 --iters 2
CNT=4 run blur 'blur' blur --pairs 4 --iters 1
CNT=2 run sharpen 'sharpen|fx_tile' sharpen --pairs 4 --iters 1
CNT=2 run adaptive 'sharpen|fx_tile' adaptive --pairs 4 --iters 1
CNT=6 run lanczos 'resize|lanczos' lanczos --pairs 2 --iters 1 --w 7680 --h 4320
CNT=14 run msssim 'box|ssim' msssim --pairs 2 --iters 1 --w 7680 --h 4320
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/bench_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
ls -la gpurun_out
