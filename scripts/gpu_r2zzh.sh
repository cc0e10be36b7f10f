#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -q > $OUT/r02zzh_pytest_gpu.log 2>&1; tail -2 $OUT/r02zzh_pytest_gpu.log | cut -c1-200
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/r02zzh_smoke.log 2>&1; tail -2 $OUT/r02zzh_smoke.log | cut -c1-250
timeout 900 python bench.py > $OUT/r02zzh_bench.json 2> $OUT/r02zzh_bench.err; python -c "
import json; d=json.load(open('$OUT/r02zzh_bench.json')); print(d['value'], d['e2e']['value'], d['ms_per_step'], d['stage_ms_per_step'], d['e2e']['step_wall_ms'], d['latency_ms'], d['cpu_baseline']['value'])"; tail -2 $OUT/r02zzh_bench.err
