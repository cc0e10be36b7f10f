#!/bin/bash
# round 2, call V: lane width of the general kernel on the massive-neutrino workloads; 200-register build of the register kernel
OUT=gpurun_out; mkdir -p $OUT
bash scripts/variants_bench.sh r02v_variants
: > $OUT/r02v_lanes.log
for w in config5_mnu_w0wa config1_massive; do for L in 8 16 32; do
  CAMB_B200_LANES=$L timeout 600 python bench.py --steps 2 --warmup 3 --workload $w --no-latency --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print('$w lanes $L', d['value'], d['stage_ms_per_step'])" >> $OUT/r02v_lanes.log 2>&1
done; done
cat $OUT/r02v_lanes.log
