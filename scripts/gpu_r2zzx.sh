#!/bin/bash
# ncu evidence at the final head of round 2: launch list of the plain bench command, full capture of the three evolve launches of one step
OUT=gpurun_out; mkdir -p $OUT
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/r02zzx_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-latency > $OUT/r02zzx_ncu_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"cb_evolve_reg" -s 9 -c 3 \
    -o $OUT/r02zzx_evolve -f python bench.py --steps 1 --warmup 3 --batch 128 --no-cpu-baseline --no-latency > $OUT/r02zzx_ncu_evolve.log 2>&1; tail -1 $OUT/r02zzx_ncu_evolve.log | cut -c1-200
