#!/bin/bash
# final state of round 2: GPU suite, smoke, default bench line, reference arm
OUT=gpurun_out; mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -q > $OUT/r02zzr_pytest_gpu.log 2>&1; tail -2 $OUT/r02zzr_pytest_gpu.log | cut -c1-200
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/r02zzr_smoke.log 2>&1; tail -1 $OUT/r02zzr_smoke.log | cut -c1-250
timeout 900 python bench.py > $OUT/r02zzr_bench.json 2> $OUT/r02zzr_bench.err; python -c "
import json; d=json.load(open('$OUT/r02zzr_bench.json')); print(d['value'], d['e2e']['value'], d['ms_per_step'], d['stage_ms_per_step'], d['e2e']['step_wall_ms'], d['roofline']['frac'], d['latency_ms']['one_planck2018_cosmology_1gpu'], d['cpu_baseline']['value'])"; tail -2 $OUT/r02zzr_bench.err
timeout 600 python bench.py --impl reference > $OUT/r02zzr_bench_reference.json 2>> $OUT/r02zzr_bench.err; python -c "
import json; d=json.load(open('$OUT/r02zzr_bench_reference.json')); print('reference', d['value'], d['cpu_baseline'])"
