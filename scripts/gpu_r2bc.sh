#!/bin/bash
# round 2, call BC: 218 registers / 4 CTAs per SM against 200 registers / 5 CTAs per SM in the many-wave regime (128 and 192 sets)
OUT=gpurun_out; mkdir -p $OUT
: > $OUT/r02bc_r200.log
for b in 128 192; do for lib in camb_b200/libcambb200.so variants/libcambb200_r200.so; do
  echo "== $lib batch $b" >> $OUT/r02bc_r200.log
  CAMB_B200_LIB=$PWD/$lib timeout 300 python scripts/evolve_only.py $b 2>&1 | tail -1 >> $OUT/r02bc_r200.log
done; done
cat $OUT/r02bc_r200.log
