#!/bin/bash
# usage: tools/build_variant.sh NAME FILE.cu "-DMACRO=V ..."   → fennec_b200/_variants/libfennec_NAME.so (FILE.cu rebuilt with the
# macros, every other object taken from the regular build).  Select it with FB_LIB_PATH=fennec_b200/_variants/libfennec_NAME.so.
set -e
name=$1; file=$2; defs=$3
root=$(cd "$(dirname "$0")/.." && pwd)
mkdir -p $root/fennec_b200/_variants
obj=$root/fennec_b200/_variants/${name}_${file%.cu}.o
/usr/local/cuda/bin/nvcc $defs -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -fmad=false \
  -Xcompiler -fPIC,-O2,-Wall,-fvisibility=hidden --expt-relaxed-constexpr -c $root/fennec_b200/csrc/$file -o $obj
objs=""
for o in $root/fennec_b200/_obj/*.o; do
  if [ "$(basename $o)" = "${file%.cu}.o" ]; then objs="$objs $obj"; else objs="$objs $o"; fi
done
/usr/local/cuda/bin/nvcc -shared -gencode arch=compute_100a,code=sm_100a -o $root/fennec_b200/_variants/libfennec_$name.so $objs -Xcompiler -fPIC -cudart static
echo built fennec_b200/_variants/libfennec_$name.so
