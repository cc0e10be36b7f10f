#!/bin/bash
# round 2, call C: launch list (per-kernel device times, serialised) and a full capture of the register kernels
OUT=gpurun_out; mkdir -p $OUT
LIB=${1:-variants/libcambb200_mb6.so}
export CAMB_B200_LIB=$PWD/$LIB
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $OUT/r02c_launches.csv \
    python bench.py --steps 1 --warmup 3 --batch 64 --no-cpu-baseline > $OUT/r02c_ncu_bench.log 2>&1
grep -E "cb_evolve" $OUT/r02c_launches.csv | tail -8
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"cb_evolve_reg" -s 6 -c 2 \
    -o $OUT/r02c_prof -f python bench.py --steps 1 --warmup 3 --batch 64 --no-cpu-baseline > $OUT/r02c_ncu_full.log 2>&1
tail -3 $OUT/r02c_ncu_full.log
ls -la $OUT/r02c*
