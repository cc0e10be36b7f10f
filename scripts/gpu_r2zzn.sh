#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
for b in 256 512; do
t0=$(date +%s)
timeout 900 python bench.py --steps 5 --warmup 3 --batch $b --no-latency --no-cpu-baseline > $OUT/r02zzn_bench_b$b.json 2> $OUT/r02zzn_bench_b$b.err
echo "batch $b wall $(( $(date +%s) - t0 )) s"
python -c "
import json; d=json.load(open('$OUT/r02zzn_bench_b$b.json')); print(d['value'], d['e2e']['value'], d['ms_per_step'], d['stage_ms_per_step'], d['e2e']['step_wall_ms'])"; tail -1 $OUT/r02zzn_bench_b$b.err
done
