#!/bin/bash
# round 2, call AC: Bessel table and thermo tables built on the device: parity; full GPU suite; bench (e2e with device-derived thermo tables)
OUT=gpurun_out; mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "bessel_table or thermo_tables" > $OUT/r02ac_new.log 2>&1; tail -15 $OUT/r02ac_new.log
timeout 1500 python -m pytest tests -m gpu -q > $OUT/r02ac_pytest_gpu.log 2>&1; echo "pytest exit $?" >> $OUT/r02ac_pytest_gpu.log
tail -4 $OUT/r02ac_pytest_gpu.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-latency --no-cpu-baseline > $OUT/r02ac_bench.json 2> $OUT/r02ac_bench.err; tail -c 1500 $OUT/r02ac_bench.json; tail -3 $OUT/r02ac_bench.err
timeout 600 python bench.py --steps 5 --warmup 3 --no-latency --no-cpu-baseline --host-thermo-tables 2>/dev/null | tail -1 | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print('host thermo tables: value', d['value'], 'e2e', d['e2e'])"
