#!/bin/bash
# Throughput against the number of parameter sets per step (wave quantisation of the persistent evolve kernel).
out=gpurun_out/${1:-batch}.log
: > $out
for b in 32 48 56 60 64 72 96 128; do
  echo "== batch $b" >> $out
  timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --batch $b 2>gpurun_out/batch_err.txt | tail -1 | python -c "
import sys, json
try:
    d = json.loads(sys.stdin.read().strip()); print(json.dumps({'value': d['value'], 'e2e': d['e2e']['value'], 'stage_ms': d['stage_ms_per_step']}))
except Exception as e: print('failed', e)" >> $out
done
cat $out
