#!/bin/bash
# palette (f3): parity, timing, and the launch list (ncu time-only pass) of the cell-list kernels
timeout 300 python -m pytest tests/test_palette.py -x -q -m gpu 2>&1 | tail -2
timeout 100 python tools/bench_ops.py palette 2>&1 | tail -2 | cut -c1-170
mkdir -p gpurun_out
timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"palette" -c 40 --csv --log-file gpurun_out/palette_launches.csv python tools/bench_ops.py palette > /dev/null 2>&1
tail -4 gpurun_out/palette_launches.csv | cut -c1-200
