#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
for ord in "1,0,2" "0,2,1" "0,1,2" "2,0,1" "1,2,0" "2,1,0"; do
  echo "order $ord (0 = Layout<4,18,31>, 1 = Layout<8,69,84>, 2 = Layout<4,29,34>)"
  CAMB_B200_LIB=variants/lib_tune.so CAMB_B200_REG_ORDER=$ord python scripts/evolve_only.py 128 2>&1 | tail -1
  CAMB_B200_LIB=variants/lib_tune.so CAMB_B200_REG_ORDER=$ord python scripts/evolve_only.py 64 2>&1 | tail -1
done > $OUT/r02zzv_class_order.log 2>&1; cat $OUT/r02zzv_class_order.log
