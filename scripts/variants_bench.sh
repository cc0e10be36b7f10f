#!/bin/bash
# Kernel-tuning experiment: bench every alternative build of libcambb200.so under variants/ (see DESIGN.md section 4.1).
out=gpurun_out/${1:-variants}.log
: > $out
for lib in camb_b200/libcambb200.so variants/*.so; do
  echo "== $lib" >> $out
  CAMB_B200_LIB=$PWD/$lib timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import sys, json
l = sys.stdin.read().strip()
try:
    d = json.loads(l); print(json.dumps({k: d[k] for k in ('value', 'ms_per_step', 'stage_ms_per_step')} | {'e2e': d['e2e']['value']}))
except Exception as e:
    print('FAILED', l[-300:])
" >> $out
done
cat $out
