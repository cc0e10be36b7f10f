#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -q -x > $OUT/r02t_pytest_gpu.log 2>&1; echo "pytest exit $?" >> $OUT/r02t_pytest_gpu.log
tail -6 $OUT/r02t_pytest_gpu.log
for w in config1_massive config5_mnu_w0wa; do
  timeout 600 python bench.py --steps 3 --warmup 3 --workload $w --no-latency --no-cpu-baseline > $OUT/r02t_bench_$w.json 2> $OUT/r02t_bench_$w.err
  python -c "
import json; d=json.load(open('$OUT/r02t_bench_$w.json')); print('$w', d['value'], d['e2e']['value'], d['stage_ms_per_step'])" 2>&1 | tail -1
done
