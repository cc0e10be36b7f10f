#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
timeout 600 compute-sanitizer --tool memcheck --print-limit 20 python scripts/debug_mpl.py > $OUT/r02l_memcheck.log 2>&1
grep -E "ERROR SUMMARY|Invalid|=========.*at |max rel|deviating" $OUT/r02l_memcheck.log | head -20
timeout 600 compute-sanitizer --tool initcheck --print-limit 20 python scripts/debug_mpl.py > $OUT/r02l_initcheck.log 2>&1
grep -E "ERROR SUMMARY|Uninit|=========.*at |max rel" $OUT/r02l_initcheck.log | head -20
timeout 600 compute-sanitizer --tool racecheck --print-limit 20 python scripts/debug_mpl.py > $OUT/r02l_racecheck.log 2>&1
grep -E "RACECHECK SUMMARY|hazard|=========.*at |max rel" $OUT/r02l_racecheck.log | head -20
