#!/bin/bash
# round 2, call AA: does the register kernel's throughput follow the number of resident blocks?  (one layout class alone)
OUT=gpurun_out; mkdir -p $OUT
: > $OUT/r02aa_residency.log
for cls in 6 8; do for nb in 1 2 3 4; do
  echo "== class $cls, $nb blocks/SM (218 registers)" >> $OUT/r02aa_residency.log
  CAMB_B200_LIB=$PWD/variants/libcambb200_tune.so CAMB_B200_ONLY_CLASS=$cls CAMB_B200_REG_BLOCKS=$nb timeout 300 python scripts/evolve_only.py 64 2>&1 | tail -1 >> $OUT/r02aa_residency.log
done
  echo "== class $cls, 5 blocks/SM (200 registers)" >> $OUT/r02aa_residency.log
  CAMB_B200_LIB=$PWD/variants/libcambb200_tune200.so CAMB_B200_ONLY_CLASS=$cls timeout 300 python scripts/evolve_only.py 64 2>&1 | tail -1 >> $OUT/r02aa_residency.log
  echo "== class $cls, 4 blocks/SM (200 registers)" >> $OUT/r02aa_residency.log
  CAMB_B200_LIB=$PWD/variants/libcambb200_tune200.so CAMB_B200_ONLY_CLASS=$cls CAMB_B200_REG_BLOCKS=4 timeout 300 python scripts/evolve_only.py 64 2>&1 | tail -1 >> $OUT/r02aa_residency.log
done
for b in 32 128; do
  echo "== class 6, 4 blocks/SM, batch $b" >> $OUT/r02aa_residency.log
  CAMB_B200_LIB=$PWD/variants/libcambb200_tune.so CAMB_B200_ONLY_CLASS=6 timeout 300 python scripts/evolve_only.py $b 2>&1 | tail -1 >> $OUT/r02aa_residency.log
done
cat $OUT/r02aa_residency.log
