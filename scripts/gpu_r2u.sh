#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"cb_evolve_kernel" -s 4 -c 2 \
    -o $OUT/r02u_prof -f python bench.py --steps 1 --warmup 3 --batch 32 --workload config5_mnu_w0wa --no-cpu-baseline --no-latency > $OUT/r02u_ncu_full.log 2>&1
tail -2 $OUT/r02u_ncu_full.log
CAMB_B200_REG=0 timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-latency 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('general kernel, lean batch 64', d['value'], d['stage_ms_per_step'])"
