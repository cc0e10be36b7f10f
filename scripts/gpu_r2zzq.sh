#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
for rep in 1 2; do python scripts/evolve_only.py 128 2>&1 | tail -1; python scripts/evolve_only.py 64 2>&1 | tail -1; done > $OUT/r02zzq_trptr.log 2>&1; cat $OUT/r02zzq_trptr.log
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_sigma8_anchor.py -m gpu -q 2>&1 | tail -2
