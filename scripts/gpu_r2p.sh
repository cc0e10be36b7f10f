#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"cb_evolve_mpl" -s 3 -c 1 \
    -o $OUT/r02p_prof -f python bench.py --steps 1 --warmup 3 --batch 64 --no-cpu-baseline --no-latency > $OUT/r02p_ncu_full.log 2>&1
tail -2 $OUT/r02p_ncu_full.log
