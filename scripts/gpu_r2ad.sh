#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
timeout 300 python scripts/debug_bessel_build.py > $OUT/r02ad_bessel.log 2>&1; cat $OUT/r02ad_bessel.log | cut -c1-300
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "thermo_tables" 2>&1 | tail -25 | cut -c1-300
