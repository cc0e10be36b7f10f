#!/bin/bash
# round 2, call W: output branch fused into the seventh stage evaluation (register and general kernels): parity + bench
OUT=gpurun_out; mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -q -x > $OUT/r02w_pytest_gpu.log 2>&1; echo "pytest exit $?" >> $OUT/r02w_pytest_gpu.log
tail -4 $OUT/r02w_pytest_gpu.log
: > $OUT/r02w_bench.log
for w in lcdm_massless config1_massive config5_mnu_w0wa; do
  timeout 600 python bench.py --steps 3 --warmup 3 --workload $w --no-latency --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print('$w', d['value'], d['e2e']['value'], d['stage_ms_per_step'])" >> $OUT/r02w_bench.log 2>&1
done
CAMB_B200_LANES=32 timeout 600 python bench.py --steps 3 --warmup 3 --workload config5_mnu_w0wa --no-latency --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print('config5 lanes 32', d['value'], d['e2e']['value'], d['stage_ms_per_step'])" >> $OUT/r02w_bench.log 2>&1
cat $OUT/r02w_bench.log
