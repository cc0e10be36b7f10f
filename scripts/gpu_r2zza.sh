#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -q > $OUT/r02zza_pytest_gpu.log 2>&1; tail -3 $OUT/r02zza_pytest_gpu.log | cut -c1-300
timeout 900 python bench.py --steps 5 --warmup 3 > $OUT/r02zza_bench.json 2> $OUT/r02zza_bench.err; python -c "
import json; d=json.load(open('$OUT/r02zza_bench.json')); print(d['value'], d['e2e']['value'], d['ms_per_step'], d['stage_ms_per_step'], d['roofline'], d.get('latency_ms'))"; tail -2 $OUT/r02zza_bench.err
timeout 600 python bench.py --steps 5 --warmup 3 --batch 64 --no-latency --no-cpu-baseline > $OUT/r02zza_bench_b64.json 2>> $OUT/r02zza_bench.err; python -c "
import json; d=json.load(open('$OUT/r02zza_bench_b64.json')); print(d['value'], d['e2e']['value'], d['ms_per_step'], d['stage_ms_per_step'])"
