#!/bin/bash
# final state of round 2 (class launch order C, A, B): GPU suite, default bench line
OUT=gpurun_out; mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -q > $OUT/r02zzw_pytest_gpu.log 2>&1; tail -2 $OUT/r02zzw_pytest_gpu.log | cut -c1-200
timeout 900 python bench.py > $OUT/r02zzw_bench.json 2> $OUT/r02zzw_bench.err; python -c "
import json; d=json.load(open('$OUT/r02zzw_bench.json')); print(d['value'], d['e2e']['value'], d['ms_per_step'], d['stage_ms_per_step'], d['e2e']['step_wall_ms'], d['roofline']['frac'], d['latency_ms']['one_planck2018_cosmology_1gpu'], d['cpu_baseline']['value'])"; tail -2 $OUT/r02zzw_bench.err
timeout 300 python bench.py --batch 64 --no-latency --no-cpu-baseline > $OUT/r02zzw_bench_b64.json 2>> $OUT/r02zzw_bench.err; python -c "
import json; d=json.load(open('$OUT/r02zzw_bench_b64.json')); print(d['value'], d['e2e']['value'], d['ms_per_step'], d['stage_ms_per_step']['evolve'], d['roofline']['frac'])"
