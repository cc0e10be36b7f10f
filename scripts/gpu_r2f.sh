#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
bash scripts/variants_bench.sh r02f_variants
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"cb_evolve_reg" -s 9 -c 3 \
    -o $OUT/r02f_prof -f python bench.py --steps 1 --warmup 3 --batch 64 --no-cpu-baseline > $OUT/r02f_ncu_full.log 2>&1
tail -2 $OUT/r02f_ncu_full.log
