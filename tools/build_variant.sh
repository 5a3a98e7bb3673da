#!/bin/bash
# usage: tools/build_variant.sh name [nvcc flags ...]: builds variants/libkpeg_cuda_<name>.so from the working tree
# (kernels.cu + kpeg_cuda.cu compiled with the extra flags) for A/B runs with KPEG_CUDA_LIB=variants/...
set -e
cd "$(dirname "$0")/.."
name=$1; shift
mkdir -p variants/obj_$name
F="-std=c++17 -O3 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC -Iinclude -Ilibkpeg_b200/csrc $*"
nvcc $F -c libkpeg_b200/csrc/kernels.cu -o variants/obj_$name/kernels.o &
nvcc $F -c libkpeg_b200/csrc/kpeg_cuda.cu -o variants/obj_$name/kpeg_cuda.o &
wait
g++ -std=c++17 -O2 -fPIC -Iinclude -c libkpeg_b200/csrc/jfif_parse.cpp -o variants/obj_$name/jfif_parse.o
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o variants/libkpeg_cuda_$name.so variants/obj_$name/*.o
echo built variants/libkpeg_cuda_$name.so
