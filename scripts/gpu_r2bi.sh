#!/bin/bash
# round 2, call BI: line-of-sight kernel with the loads of two (1, 3, 4) time steps issued ahead of the arithmetic; cold evolve code out of line
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -q -x > $OUT/r02bi_pytest_gpu.log 2>&1; tail -2 $OUT/r02bi_pytest_gpu.log
: > $OUT/r02bi_los.log
for lib in camb_b200/libcambb200.so variants/libcambb200_los_u1.so variants/libcambb200_los_u3.so variants/libcambb200_los_u4.so variants/libcambb200_coldnoinl.so; do
  echo "== $lib" >> $OUT/r02bi_los.log
  CAMB_B200_LIB=$PWD/$lib timeout 600 python bench.py --steps 3 --warmup 3 --no-latency --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['e2e']['value'], d['stage_ms_per_step'])" >> $OUT/r02bi_los.log 2>&1
done
cat $OUT/r02bi_los.log
