#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
for rep in 1 2; do for lib in "" variants/lib_skip.so; do echo "lib=$lib"; CAMB_B200_LIB=$lib python scripts/evolve_only.py 128 2>&1 | tail -1; CAMB_B200_LIB=$lib python scripts/evolve_only.py 64 2>&1 | tail -1; done; done > $OUT/r02zzs_skip_ab.log 2>&1; cat $OUT/r02zzs_skip_ab.log
CAMB_B200_LIB=variants/lib_skip.so timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -2
