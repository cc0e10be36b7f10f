#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
timeout 600 python bench.py --steps 5 --warmup 3 --no-latency --no-cpu-baseline > $OUT/r02ah_bench.json 2> $OUT/r02ah_bench.err; python -c "
import json; d=json.load(open('$OUT/r02ah_bench.json')); print(d['value'], d['e2e']['value'], d['ms_per_step'], d['stage_ms_per_step'])"; tail -3 $OUT/r02ah_bench.err
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
