#!/bin/bash
# Build libkcrn variants for same-box A/B timing (KCRN_LIBRARY=<so> selects one at run time):
#   profiles/microbench/build_variants.sh "tag:minblocks:tilennz[:-DNAME=VALUE ...]" ...
#   -> profiles/microbench/_build/libkcrn_<tag>.so        (minblocks / tilennz: '-' keeps the source's value)
set -e
ROOT=$(cd "$(dirname "$0")/../.." && pwd)
SRC=$ROOT/krylov-cubic-regularized-newton_b200/csrc
OUT=$ROOT/profiles/microbench/_build
mkdir -p $OUT
for v in "$@"; do
  IFS=: read tag mb tn defs <<< "$v"
  W=/tmp/kcrn_var_$tag; rm -rf $W; mkdir -p $W; cp $SRC/*.cu $SRC/*.cuh $W/
  [ "$mb" != "-" ] && [ -n "$mb" ] && sed -i "s/__launch_bounds__(kSweepThreads, 4) sweep_kernel/__launch_bounds__(kSweepThreads, $mb) sweep_kernel/" $W/kcrn_sweep.cu
  [ "$tn" != "-" ] && [ -n "$tn" ] && sed -i "s/constexpr int kTileNnz = 2048;/constexpr int kTileNnz = $tn;/" $W/kcrn_internal.cuh
  objs=""
  for f in $W/*.cu; do
    nvcc $defs -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -I$ROOT/include -I$W -c $f -o ${f%.cu}.o &
    objs="$objs ${f%.cu}.o"
  done
  wait
  nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $OUT/libkcrn_$tag.so $objs -ldl
  echo built $OUT/libkcrn_$tag.so
done
