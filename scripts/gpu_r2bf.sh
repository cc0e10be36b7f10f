#!/bin/bash
# round 2, call BF: instruction-cache footprint experiments on the register kernel (stage combinations as one loop; stage_values out of line)
OUT=gpurun_out; mkdir -p $OUT
: > $OUT/r02bf_icache.log
for rep in 1 2; do for lib in camb_b200/libcambb200.so variants/libcambb200_loop.so variants/libcambb200_loop_noinl.so; do
  echo "== $lib" >> $OUT/r02bf_icache.log
  CAMB_B200_LIB=$PWD/$lib timeout 300 python scripts/evolve_only.py 128 2>&1 | tail -1 >> $OUT/r02bf_icache.log
done; done
for lib in camb_b200/libcambb200.so variants/libcambb200_loop.so variants/libcambb200_loop_noinl.so; do
  echo "== ncu $lib" >> $OUT/r02bf_icache.log
  CAMB_B200_LIB=$PWD/$lib timeout 600 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,sm__icc_request_hit_rate.pct,gcc__cache_requests_type_instruction.sum.pct_of_peak_sustained_elapsed --clock-control none -k regex:"cb_evolve_reg" -s 3 -c 3 --csv python scripts/evolve_only.py 64 2>&1 | grep "cb_evolve_reg" | awk -F'","' '{print substr($5,40,36), $(NF-2), $NF}' >> $OUT/r02bf_icache.log
done
cat $OUT/r02bf_icache.log
