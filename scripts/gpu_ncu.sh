#!/bin/bash
# full ncu capture of the evolve kernel (first launch of the small-nv class) at a given lane width
OUT=gpurun_out; mkdir -p $OUT
L=${1:-8}; B=${2:-32}; TAG=${3:-exp}
CAMB_B200_LANES=$L timeout 900 ncu --set full --clock-control none --import-source on -k regex:"cb_evolve_kernel" -c 1 \
    -o $OUT/${TAG}_prof_L$L -f python bench.py --steps 1 --warmup 3 --batch $B --no-cpu-baseline > $OUT/${TAG}_ncu_L$L.log 2>&1
tail -3 $OUT/${TAG}_ncu_L$L.log
