#!/bin/bash
# round 2, call D: register kernel with the lowest-k layout C; variants (register caps, 16-lane modes through the general kernel)
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -q -x > $OUT/r02d_pytest_gpu.log 2>&1; echo "pytest exit $?" >> $OUT/r02d_pytest_gpu.log
tail -5 $OUT/r02d_pytest_gpu.log
bash scripts/variants_bench.sh r02d_variants
echo "== default lib, CAMB_B200_REG=2 (16-lane layout's modes through the general kernel)" >> $OUT/r02d_variants.log
CAMB_B200_REG=2 timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print(json.dumps({k: d[k] for k in ('value', 'ms_per_step', 'stage_ms_per_step')} | {'e2e': d['e2e']['value']}))" >> $OUT/r02d_variants.log
tail -3 $OUT/r02d_variants.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $OUT/r02d_launches.csv \
    python bench.py --steps 1 --warmup 3 --batch 64 --no-cpu-baseline > $OUT/r02d_ncu_bench.log 2>&1
grep -E "cb_evolve" $OUT/r02d_launches.csv | awk -F'","' '{print $5, $NF}' | tail -4
