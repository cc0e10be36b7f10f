#!/bin/bash
# usage: scripts/build_variant.sh <name> <file.cu> [nvcc flags...]  -> variants/libcambb200_<name>.so
# (the named source compiled with the extra flags, every other object from build/cuda as is)
name=$1; src=$2; shift 2
base=$(basename $src .cu)
mkdir -p variants build/variants
/usr/local/cuda/bin/nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC -Xptxas -v "$@" \
  -c -o build/variants/${base}_$name.o $src 2> build/variants/${base}_$name.ptxas.log || { cat build/variants/${base}_$name.ptxas.log; exit 1; }
objs=$(ls build/cuda/*.o | grep -v "/$base.o")
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o variants/libcambb200_$name.so $objs build/variants/${base}_$name.o -lcudart
grep -E "Used|spill" build/variants/${base}_$name.ptxas.log | head -4
