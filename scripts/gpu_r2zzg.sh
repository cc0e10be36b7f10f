#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
for b in 1 2 4 8 16; do for mm in 0 100000000; do echo "B=$b reg_min_modes=$mm"; CAMB_B200_REG_MIN_MODES=$mm python scripts/evolve_only.py $b 2>&1 | tail -1; done; done > $OUT/r02zzg_few_modes.log 2>&1
for b in 32 96 192 256; do python scripts/evolve_only.py $b 2>&1 | tail -1; done > $OUT/r02zzg_batch_sweep.log 2>&1
cat $OUT/r02zzg_few_modes.log $OUT/r02zzg_batch_sweep.log
