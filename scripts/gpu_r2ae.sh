#!/bin/bash
# round 2, call AE: full GPU suite (device Bessel table, device thermo packing, non-linear source scaling), bench with the e2e leg on device-derived thermo tables
OUT=gpurun_out; mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -q > $OUT/r02ae_pytest_gpu.log 2>&1; echo "pytest exit $?" >> $OUT/r02ae_pytest_gpu.log
tail -6 $OUT/r02ae_pytest_gpu.log | cut -c1-300
timeout 600 python bench.py --steps 5 --warmup 3 --no-latency --no-cpu-baseline > $OUT/r02ae_bench.json 2> $OUT/r02ae_bench.err; python -c "
import json; d=json.load(open('$OUT/r02ae_bench.json')); print(d['value'], d['e2e'], d['stage_ms_per_step'])"; tail -3 $OUT/r02ae_bench.err
timeout 600 python bench.py --steps 5 --warmup 3 --no-latency --no-cpu-baseline --host-thermo-tables 2>/dev/null | tail -1 | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print('host thermo tables: value', d['value'], 'e2e', d['e2e'])"
