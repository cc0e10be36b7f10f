#!/bin/bash
# round 2, call BE: warp-convergent derivs (full-mask shuffles) against the per-group-mask build: instruction counts (ncu) and times
OUT=gpurun_out; mkdir -p $OUT
: > $OUT/r02be_ab.log
for rep in 1 2; do for lib in camb_b200/libcambb200.so variants/libcambb200_divshfl.so; do
  echo "== $lib" >> $OUT/r02be_ab.log
  CAMB_B200_LIB=$PWD/$lib timeout 300 python scripts/evolve_only.py 128 2>&1 | tail -1 >> $OUT/r02be_ab.log
done; done
for lib in camb_b200/libcambb200.so variants/libcambb200_divshfl.so; do
  echo "== ncu $lib" >> $OUT/r02be_ab.log
  CAMB_B200_LIB=$PWD/$lib timeout 600 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio,sm__icc_request_hit_rate.pct,gcc__cache_requests_type_instruction.sum.pct_of_peak_sustained_elapsed --clock-control none -k regex:"cb_evolve_reg" -s 3 -c 3 --csv python scripts/evolve_only.py 64 2>&1 | grep "cb_evolve_reg" | awk -F'","' '{print substr($5,40,36), $(NF-2), $NF}' >> $OUT/r02be_ab.log
done
cat $OUT/r02be_ab.log
