#!/bin/bash
# round 2, call Y: line-of-sight kernel with the lanes on the multipoles (transposed Bessel table): parity, A/B bench, ncu
OUT=gpurun_out; mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -q -x > $OUT/r02y_pytest_gpu.log 2>&1; echo "pytest exit $?" >> $OUT/r02y_pytest_gpu.log
tail -4 $OUT/r02y_pytest_gpu.log
: > $OUT/r02y_bench.log
for los in 1 0; do
  CAMB_B200_LOS=$los timeout 600 python bench.py --steps 3 --warmup 3 --no-latency --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print('los_kernel $los', d['value'], d['e2e']['value'], d['stage_ms_per_step'])" >> $OUT/r02y_bench.log 2>&1
done
for w in config1_massive single_cosmology; do
  timeout 600 python bench.py --steps 3 --warmup 3 --workload $w --no-latency --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print('$w', d['value'], d['e2e']['value'], d['stage_ms_per_step'])" >> $OUT/r02y_bench.log 2>&1
done
cat $OUT/r02y_bench.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"cb_los_kernel" -s 3 -c 1 \
    -o $OUT/r02y_los -f python bench.py --steps 1 --warmup 3 --batch 64 --no-cpu-baseline --no-latency > $OUT/r02y_ncu_full.log 2>&1
tail -1 $OUT/r02y_ncu_full.log
