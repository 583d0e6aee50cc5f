#!/bin/bash
# Build alternative libraries for A/B timing on the GPU box (run HERE, before gpurun: nvcc cross-compiles without a GPU
# and build/ travels with the snapshot).  Usage:
#   bash tools/build_variants.sh modes                 -> build/dbg/lib_m{1,2,4,6,7,8,14}.so   (TC_DEBUG_MODE, tools/dbg_modes.sh)
#   bash tools/build_variants.sh tag "-DFOO=1 -DBAR=2" -> build/dbg/lib_tag.so                 (tools/ab_libs.sh)
set -e
cd "$(dirname "$0")/.."
mkdir -p build/dbg
NVCC="nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -shared -Xcompiler -fPIC -I include -I implicit_depth_b200/csrc"
if [ "$1" = "modes" ]; then
  for m in 1 2 4 6 7 8 14; do $NVCC -DTC_DEBUG_MODE=$m -o build/dbg/lib_m$m.so implicit_depth_b200/csrc/lidf_query.cu & done
  wait
else
  $NVCC $2 -o build/dbg/lib_$1.so implicit_depth_b200/csrc/lidf_query.cu
fi
ls -la build/dbg
