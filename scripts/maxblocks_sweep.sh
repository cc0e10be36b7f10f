#!/bin/bash
# Does evolve throughput scale with the number of modes in flight per SM?  (CAMB_B200_MAXBLOCKS caps resident blocks.)
out=gpurun_out/${1:-maxblocks}.log
: > $out
for mb in 2 3 4 5 6; do
  echo "== maxblocks $mb" >> $out
  CAMB_B200_MAXBLOCKS=$mb timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip()); print(json.dumps({'value': d['value'], 'stage_ms': d['stage_ms_per_step']}))" >> $out
done
cat $out
