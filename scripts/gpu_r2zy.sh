#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
for rep in 1 2; do
for lib in variants/lib_w8.so variants/lib_ns12w2.so variants/lib_ns12w6.so variants/lib_ns12w7.so; do
  echo "lib=$lib"; CAMB_B200_LIB=$lib timeout 300 python scripts/evolve_only.py 128 2>&1 | tail -1
  [ $rep = 1 ] && CAMB_B200_LIB=$lib timeout 300 python scripts/evolve_only.py 64 2>&1 | tail -1
done; done > $OUT/r02zy_ns12_ab.log 2>&1
cat $OUT/r02zy_ns12_ab.log
CAMB_B200_LIB=variants/lib_ns12w7.so timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "kernel_variant or step_stat or parity" 2>&1 | tail -5
