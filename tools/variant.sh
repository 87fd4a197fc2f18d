#!/bin/bash
# development: build a variant of one translation unit and link it with the objects of the regular build
# usage: tools/variant.sh <name> <file.cu> <extra nvcc flags...>   ->  psl_slam_b200/libpsl_frontend_<name>.so
set -e
cd "$(dirname "$0")/../psl_slam_b200/csrc"
name=$1; src=$2; shift 2
mkdir -p build_var
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -fmad=false -Xcompiler -fPIC -Xptxas -v "$@" -c $src -o build_var/${name}.o 2> build_var/${name}.log
objs=$(ls build/*.o | grep -v "build/${src%.cu}.o")
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../libpsl_frontend_${name}.so $objs build_var/${name}.o -lcudart
grep -A2 "lsd_core_kernel" build_var/${name}.log | grep -E "registers|spill" | tr '\n' ' '; echo
