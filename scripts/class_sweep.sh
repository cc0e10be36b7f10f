#!/bin/bash
# needs a library built with -DCB_TUNING_KNOBS (make cuda NVFLAGS+=-DCB_TUNING_KNOBS)
# Duration of every slab class of the evolve stage when it runs alone (bench batch), at 6 and 3 resident blocks per SM.
out=gpurun_out/${1:-classes}.log
: > $out
for mb in 6 3; do for c in 0 1 2 3 4; do
  echo "== class $c maxblocks $mb" >> $out
  CAMB_B200_ONLY_CLASS=$c CAMB_B200_MAXBLOCKS=$mb timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>gpurun_out/class_err.txt | tail -1 | python -c "
import sys, json
try:
    d = json.loads(sys.stdin.read().strip()); print(json.dumps({'evolve_ms': d['stage_ms_per_step']['evolve']}))
except Exception as e: print('failed', e)" >> $out
done; done
cat $out
