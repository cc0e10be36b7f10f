#!/bin/bash
# round 2, call ZZK: test suite, bench lines of every workload, and the ncu evidence at the head (12 elements per lane, slim slabs, fused output,
# division fix)
OUT=gpurun_out; mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -q > $OUT/r02zzk_pytest_gpu.log 2>&1; tail -2 $OUT/r02zzk_pytest_gpu.log | cut -c1-200
timeout 900 python bench.py --steps 5 --warmup 3 > $OUT/r02zzk_bench.json 2> $OUT/r02zzk_bench.err
timeout 600 python bench.py --steps 5 --warmup 3 --batch 64 --no-latency --no-cpu-baseline > $OUT/r02zzk_bench_b64.json 2>> $OUT/r02zzk_bench.err
timeout 900 python bench.py --steps 3 --warmup 3 --workload config1_massive > $OUT/r02zzk_bench_config1.json 2>> $OUT/r02zzk_bench.err
timeout 900 python bench.py --steps 3 --warmup 3 --workload config5_mnu_w0wa > $OUT/r02zzk_bench_config5.json 2>> $OUT/r02zzk_bench.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/r02zzk_bench_reference.json 2>> $OUT/r02zzk_bench.err
for f in bench bench_b64 bench_config1 bench_config5 bench_reference; do python -c "
import json; d=json.load(open('$OUT/r02zzk_$f.json')); print('$f', round(d['value'],1), round(d['e2e']['value'],1), d.get('ms_per_step'), d.get('stage_ms_per_step'), (d.get('roofline') or {}).get('frac'))"; done
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/r02zzk_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-latency > $OUT/r02zzk_ncu_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"cb_evolve_reg" -s 9 -c 3 \
    -o $OUT/r02zzk_evolve -f python bench.py --steps 1 --warmup 3 --batch 128 --no-cpu-baseline --no-latency > $OUT/r02zzk_ncu_evolve.log 2>&1; tail -1 $OUT/r02zzk_ncu_evolve.log | cut -c1-200
