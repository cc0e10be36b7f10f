#!/bin/bash
# round 2, call Z: GPU tests after the LOS slot fix; lane width per slab-size class on the massive-neutrino workloads
OUT=gpurun_out; mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -q -x > $OUT/r02z_pytest_gpu.log 2>&1; echo "pytest exit $?" >> $OUT/r02z_pytest_gpu.log
tail -4 $OUT/r02z_pytest_gpu.log
: > $OUT/r02z_lanes.log
for w in config5_mnu_w0wa config1_massive; do
  CAMB_B200_VERBOSE=1 timeout 600 python bench.py --steps 1 --warmup 3 --workload $w --no-latency --no-cpu-baseline 2>&1 | grep "cb_evolve: class" | sort | uniq -c >> $OUT/r02z_lanes.log
  for CL in 8,8,8,8,32 8,8,8,16,32 8,8,16,32,32 8,16,16,32,32 16,16,16,32,32 8,8,8,32,32; do
  CAMB_B200_CLASS_LANES=$CL timeout 600 python bench.py --steps 2 --warmup 3 --workload $w --no-latency --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print('$w class lanes $CL', d['value'], d['stage_ms_per_step']['evolve'])" >> $OUT/r02z_lanes.log 2>&1
done; done
cat $OUT/r02z_lanes.log
