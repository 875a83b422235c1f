#!/bin/bash
# build_variant.sh SUFFIX [extra nvcc flags...] -- a second copy of the library with other compile-time settings, for A/B
# runs on the GPU box in one call (DRMLT_B200_LIB=.../libdrmlt_b200_SUFFIX.so selects it; tuning aid, not a product path)
set -e
cd "$(dirname "$0")/../drmlt-mitsuba_b200/csrc"
suffix=$1; shift
mkdir -p build_$suffix
for f in drmlt_b200.cu k_chain.cu k_walk.cu k_pt.cu k_bdpt.cu k_direct.cu k_trace.cu k_util.cu multi_gpu.cu bvh_gpu.cu bvh_build.cpp; do
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC "$@" -c $f -o build_$suffix/${f%.*}.o &
done
wait
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o libdrmlt_b200_$suffix.so build_$suffix/*.o
echo built libdrmlt_b200_$suffix.so
