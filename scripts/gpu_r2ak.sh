#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
ASYNC_OUT=1 timeout 300 python scripts/e2e_gaps.py 2>&1 | tail -2
timeout 600 python bench.py --steps 6 --warmup 3 --no-latency --no-cpu-baseline > $OUT/r02ak_bench.json 2> $OUT/r02ak_bench.err; python -c "
import json; d=json.load(open('$OUT/r02ak_bench.json')); print(d['value'], d['e2e'], d['ms_per_step'], d['stage_ms_per_step'])"; tail -3 $OUT/r02ak_bench.err
