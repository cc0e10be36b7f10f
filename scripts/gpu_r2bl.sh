#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -q -k "halofit or nonlinear" > $OUT/r02bl_nonlinear.log 2>&1; tail -5 $OUT/r02bl_nonlinear.log | cut -c1-300; grep -h "P_nl\|non-linear lensing chain" $OUT/r02bl_nonlinear.log | head
timeout 900 python -m pytest tests -m gpu -q -s -k "halofit or nonlinear_lensing_chain" 2>&1 | grep -E "P_nl|non-linear lensing chain|passed|failed" | cut -c1-300
