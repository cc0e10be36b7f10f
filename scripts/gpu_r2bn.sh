#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
: > $OUT/r02bn_los_prefetch.log
for lib in camb_b200/libcambb200.so variants/libcambb200_los_pf1.so variants/libcambb200_los_pf2.so variants/libcambb200_los_pf4.so; do
  echo "== $lib" >> $OUT/r02bn_los_prefetch.log
  CAMB_B200_LIB=$PWD/$lib timeout 600 python bench.py --steps 3 --warmup 3 --no-latency --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['e2e']['value'], d['stage_ms_per_step'])" >> $OUT/r02bn_los_prefetch.log 2>&1
done
cat $OUT/r02bn_los_prefetch.log
