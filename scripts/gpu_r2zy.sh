#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
for w in config1_massive config5_mnu_w0wa; do
  timeout 900 python bench.py --steps 3 --warmup 3 --workload $w --no-latency > $OUT/r02zz_bench_$w.json 2> $OUT/r02zz_bench_$w.err; echo "bench $w exit $?"
  python -c "
import json; d=json.load(open('$OUT/r02zz_bench_$w.json')); print('$w', d['config']['sets_per_gpu_per_step'], d['value'], d['e2e']['value'], d['stage_ms_per_step'], d.get('cpu_baseline',{}).get('value'), d['roofline']['frac'])" 2>&1 | tail -1
done
