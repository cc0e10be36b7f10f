#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/exp_pytest.log 2>&1; tail -2 $OUT/exp_pytest.log
run() { name=$1; shift
  env "$@" > $OUT/exp_bench_$name.json 2> $OUT/exp_bench_$name.err
  python - <<PY
import json
try:
    d=json.load(open("$OUT/exp_bench_$name.json")); print("$name", "value %.1f e2e %.1f"%(d["value"], d["e2e"]["value"]), d["stage_ms_per_step"])
except Exception as e: print("$name failed", e); print(open("$OUT/exp_bench_$name.err").read()[-1500:])
PY
}
run b64 timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --batch 64
run b128 timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --batch 128
