#!/bin/bash
# round 2, call B: parity of the register-resident evolve kernel + A/B of its register caps against the round-1 library
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -q > $OUT/r02b_pytest_gpu.log 2>&1; echo "pytest exit $?" >> $OUT/r02b_pytest_gpu.log
tail -15 $OUT/r02b_pytest_gpu.log
bash scripts/variants_bench.sh r02b_variants
