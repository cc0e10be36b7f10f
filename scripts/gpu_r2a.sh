#!/bin/bash
# round 2, call A: parity of the fused-output evolve kernel + A/B against the round-1 library
OUT=gpurun_out; mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > $OUT/r02a_smi.txt 2>&1; nproc >> $OUT/r02a_smi.txt
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/r02a_pytest_gpu.log 2>&1; echo "pytest exit $?" >> $OUT/r02a_pytest_gpu.log
tail -4 $OUT/r02a_pytest_gpu.log
bash scripts/variants_bench.sh r02a_variants
