#!/bin/bash
# full ncu captures at the bench batch size: first evolve launch (small-nv class, 8 lanes per mode) and the LOS kernel
OUT=gpurun_out; mkdir -p $OUT
TAG=${1:-r01c}; B=${2:-64}
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"cb_evolve_kernel" -s 6 -c 2 \
    -o $OUT/${TAG}_evolve -f python bench.py --steps 1 --warmup 3 --batch $B --no-cpu-baseline > $OUT/${TAG}_ncu_evolve.log 2>&1
tail -2 $OUT/${TAG}_ncu_evolve.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"cb_los_kernel" -s 3 -c 1 \
    -o $OUT/${TAG}_los -f python bench.py --steps 1 --warmup 3 --batch $B --no-cpu-baseline > $OUT/${TAG}_ncu_los.log 2>&1
tail -2 $OUT/${TAG}_ncu_los.log
