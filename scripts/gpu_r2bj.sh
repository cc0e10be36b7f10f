#!/bin/bash
# round 2, call BJ: line-of-sight rows by bulk copies (TMA) into per-warp shared-memory rings against the LDG rows and the gather kernel
OUT=gpurun_out; mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "los_and_cl_stage or cl_parity_full_chain or tensor_cl or bessel_table or asynchronous or config3" > $OUT/r02bj_pytest_los.log 2>&1; tail -3 $OUT/r02bj_pytest_los.log | cut -c1-300
: > $OUT/r02bj_los.log
for los in 2 1 0; do
  echo "== los_kernel $los" >> $OUT/r02bj_los.log
  CAMB_B200_LOS=$los timeout 600 python bench.py --steps 3 --warmup 3 --no-latency --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['e2e']['value'], d['stage_ms_per_step'])" >> $OUT/r02bj_los.log 2>&1
done
cat $OUT/r02bj_los.log
timeout 900 python -m pytest tests -m gpu -q > $OUT/r02bj_pytest_gpu.log 2>&1; tail -3 $OUT/r02bj_pytest_gpu.log | cut -c1-300
