#!/bin/bash
# round 2, call BB: launch order of the three register-layout classes (0 = q >= 0.05 bulk, 1 = 16-lane deep low-k, 2 = lowest k)
OUT=gpurun_out; mkdir -p $OUT
: > $OUT/r02bb_order.log
for ord in 1,0,2 0,2,1 0,1,2 2,0,1 1,2,0 2,1,0; do
  echo "== register-class launch order $ord" >> $OUT/r02bb_order.log
  CAMB_B200_LIB=$PWD/variants/libcambb200_tune.so CAMB_B200_REG_ORDER=$ord timeout 300 python scripts/evolve_only.py 64 2>&1 | tail -1 >> $OUT/r02bb_order.log
done
for b in 32 96 128; do
  echo "== default order, batch $b" >> $OUT/r02bb_order.log
  CAMB_B200_LIB=$PWD/variants/libcambb200_tune.so timeout 300 python scripts/evolve_only.py $b 2>&1 | tail -1 >> $OUT/r02bb_order.log
done
cat $OUT/r02bb_order.log
