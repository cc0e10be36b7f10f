#!/bin/bash
# round 2, call E: where the state lives (y in registers or shared memory) x register cap
OUT=gpurun_out; mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_symbolic_fixture.py -m gpu -q -x > $OUT/r02e_pytest_gpu.log 2>&1; echo "pytest exit $?" >> $OUT/r02e_pytest_gpu.log
tail -3 $OUT/r02e_pytest_gpu.log
bash scripts/variants_bench.sh r02e_variants
