#!/bin/bash
# usage (on the GPU box, via gpurun): tools/gpu_sweep_r2.sh <tag> "<bench_ops ops>" [ENV=V ...] -- [variant.so ...]
# Times tools/bench_ops.py <ops> with the regular library and then with every variant library built by
# tools/build_variant.sh (fennec_b200/_variants/libfennec_<name>.so, selected through FB_LIB_PATH); extra ENV=V pairs
# apply to every run.  The lines are appended to gpurun_out/<tag>.txt — profiles/r2_tuning_sweep.txt is a digest of such runs.
tag=$1; ops=$2; shift 2
envs=()
while [ $# -gt 0 ] && [ "$1" != "--" ]; do envs+=("$1"); shift; done
[ "$1" == "--" ] && shift
out=gpurun_out/$tag.txt; mkdir -p gpurun_out
run() { echo "== $*" | tee -a $out; env "$@" 2>&1 | grep -E '"op"' | sed 's/"items_per_s.*frac_of_measured_hbm"/"frac"/' | tee -a $out; }
run "${envs[@]}" python tools/bench_ops.py $ops
for so in "$@"; do run "${envs[@]}" FB_LIB_PATH=$so python tools/bench_ops.py $ops; done
