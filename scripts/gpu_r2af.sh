#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
timeout 300 python scripts/thermo_pack_time.py 2>&1 | tail -5
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"cb_thermo" -c 4 --csv python scripts/thermo_pack_time.py 2>&1 | grep -E "cb_thermo" | awk -F'","' '{print substr($5,1,40), $NF}' | head -6
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "nonlinear or bessel_table" 2>&1 | tail -5 | cut -c1-300
