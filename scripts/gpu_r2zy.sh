#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
for rep in 1 2; do
for lib in "" variants/lib_e_ns6w2.so variants/lib_f_ns6w2.so variants/lib_g_ns6w2.so; do
  echo "lib=$lib"; CAMB_B200_LIB=$lib timeout 300 python scripts/evolve_only.py 128 2>&1 | tail -1
done; done > $OUT/r02zy_diet3_ab.log 2>&1
cat $OUT/r02zy_diet3_ab.log
