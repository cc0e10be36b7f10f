#!/bin/bash
# needs a library built with -DCB_TUNING_KNOBS (make cuda NVFLAGS+=-DCB_TUNING_KNOBS)
# Co-scheduling of the evolve slab classes: launch order (descending slab first) and blocks-per-SM caps per class.
out=gpurun_out/${1:-sched}.log
: > $out
run() {
  echo "== ORDER=$1 CAPS=$2" >> $out
  CAMB_B200_ORDER=$1 CAMB_B200_CAPS=$2 timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>gpurun_out/sched_err.txt | tail -1 | python -c "
import sys, json
try:
    d = json.loads(sys.stdin.read().strip()); print(json.dumps({'value': d['value'], 'evolve_ms': d['stage_ms_per_step']['evolve']}))
except Exception as e: print('failed', e)" >> $out
}
run 0 0,0,0,0
run 1 0,0,0,0
run 1 0,1,1,1
run 1 0,2,1,1
run 1 4,1,1,1
run 1 3,2,1,1
run 0 4,1,1,1
run 0 3,2,1,1
run 1 5,1,1,1
cat $out
