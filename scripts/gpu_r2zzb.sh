#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
for rep in 1 2; do
for lib in "" variants/lib_w4.so; do
  echo "lib=$lib"; CAMB_B200_LIB=$lib timeout 300 python scripts/evolve_only.py 128 2>&1 | tail -1
  CAMB_B200_LIB=$lib timeout 300 python scripts/evolve_only.py 64 2>&1 | tail -1
done; done > $OUT/r02zzb_w4_ab.log 2>&1
cat $OUT/r02zzb_w4_ab.log
python scripts/e2e_gaps.py 128 > gpurun_out/r02zzb_gaps.log 2>&1; NO_GC=1 python scripts/e2e_gaps.py 128 > gpurun_out/r02zzb_gaps_nogc.log 2>&1; tail -13 gpurun_out/r02zzb_gaps.log; tail -3 gpurun_out/r02zzb_gaps_nogc.log
