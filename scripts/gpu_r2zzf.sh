#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -q 2>&1 | tail -2
timeout 900 python bench.py --steps 8 --warmup 3 --no-latency --no-cpu-baseline > $OUT/r02zzf_bench.json 2> $OUT/r02zzf_bench.err; python -c "
import json; d=json.load(open('$OUT/r02zzf_bench.json')); print(d['value'], d['e2e']['value'], d['ms_per_step'], d['stage_ms_per_step'], d['e2e']['step_wall_ms'])"; tail -2 $OUT/r02zzf_bench.err
