#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
timeout 600 python scripts/debug_mpl.py 2>&1 | grep -E "max rel|deviating|stats identical"
for b in 64 192; do
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-latency --batch $b 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('mpl batch $b', d['value'], d['e2e']['value'], d['stage_ms_per_step'])"
CAMB_B200_REG=1 timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-latency --batch $b 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('reg batch $b', d['value'], d['e2e']['value'], d['stage_ms_per_step'])"
done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"cb_evolve_mpl" -s 3 -c 1 \
    -o $OUT/r02r_prof -f python bench.py --steps 1 --warmup 3 --batch 64 --no-cpu-baseline --no-latency > $OUT/r02r_ncu_full.log 2>&1
tail -1 $OUT/r02r_ncu_full.log
