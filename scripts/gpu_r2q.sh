#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
timeout 600 python scripts/debug_mpl.py 2>&1 | grep -E "max rel|deviating|stats identical"
for b in 64 128 256; do
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-latency --batch $b 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('batch $b', d['value'], d['e2e']['value'], d['stage_ms_per_step'])"
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 20 -c 30 --csv --log-file $OUT/r02q_launches.csv \
    python bench.py --steps 1 --warmup 3 --batch 64 --no-cpu-baseline --no-latency > $OUT/r02q_ncu_bench.log 2>&1
grep -E "cb_evolve" $OUT/r02q_launches.csv | awk -F'","' '{print $5, $NF}' | cut -c1-80,150- | tail -4
