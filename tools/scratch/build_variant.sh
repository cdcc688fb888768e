#!/bin/bash
# usage: build_variant.sh <name> <extra nvcc flags...>   -> tools/ab/liborbx_<name>.so
set -e
cd "$(dirname "$0")/../.."
mkdir -p tools/ab
name=$1; shift
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo --fmad=false \
  -Xcompiler -fPIC,-ffp-contract=off,-fvisibility=hidden -shared -cudart static "$@" 2>/dev/null \
  orbslam2_with_quadrics_b200/csrc/orbx_kernels.cu orbslam2_with_quadrics_b200/csrc/orbx_api.cu -o tools/ab/liborbx_$name.so
echo built tools/ab/liborbx_$name.so
