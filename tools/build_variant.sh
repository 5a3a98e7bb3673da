#!/bin/bash
# usage: tools/build_variant.sh name [nvcc flags ...]: builds variants/<name>/libkpeg_cuda.so from the working tree with the
# extra nvcc flags, for A/B runs on the GPU box with KPEG_CUDA_LIB=variants/<name>/libkpeg_cuda.so (tools/ab.sh)
set -e
cd "$(dirname "$0")/.."
name=$1; shift
make -s -C libkpeg_b200 -j8 LIBDIR=../variants/$name EXTRA="$*" ../variants/$name/libkpeg_cuda.so
echo built variants/$name/libkpeg_cuda.so
