#!/bin/bash
# round 2, call BD: asynchronous entry points + 128 sets per step
OUT=gpurun_out; mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -q > $OUT/r02bd_pytest_gpu.log 2>&1; echo "pytest exit $?" >> $OUT/r02bd_pytest_gpu.log
tail -5 $OUT/r02bd_pytest_gpu.log | cut -c1-300
for b in 128 64; do
timeout 900 python bench.py --steps 5 --warmup 3 --no-latency --no-cpu-baseline --batch $b > $OUT/r02bd_bench_b$b.json 2> $OUT/r02bd_bench_b$b.err; echo "bench exit $?"; python -c "
import json; d=json.load(open('$OUT/r02bd_bench_b$b.json')); print('batch $b', d['value'], d['e2e']['value'], d['e2e']['step_wall_ms'], d['ms_per_step'], d['stage_ms_per_step'])"; tail -3 $OUT/r02bd_bench_b$b.err
done
