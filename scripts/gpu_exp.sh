#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -x -q -s > $OUT/exp_pytest.log 2>&1; tail -4 $OUT/exp_pytest.log; grep -E "C_l errors|identical" $OUT/exp_pytest.log | head -30
run() { name=$1; shift
  env "$@" timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $OUT/exp_bench_$name.json 2> $OUT/exp_bench_$name.err
  python - <<PY
import json
try:
    d=json.load(open("$OUT/exp_bench_$name.json")); print("$name", "value %.1f e2e %.1f"%(d["value"], d["e2e"]["value"]), d["stage_ms_per_step"])
except Exception as e: print("$name failed", e); print(open("$OUT/exp_bench_$name.err").read()[-1500:])
PY
}
run L8 CAMB_B200_LANES=8
run L16 CAMB_B200_LANES=16
run L32 CAMB_B200_LANES=32
