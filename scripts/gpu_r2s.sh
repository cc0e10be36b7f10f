#!/bin/bash
# round 2, call S: full GPU test-suite, smoke, bench lines of every workload, reference arm
OUT=gpurun_out; mkdir -p $OUT
timeout 1200 python -m pytest tests -m gpu -q > $OUT/r02s_pytest_gpu.log 2>&1; echo "pytest exit $?" >> $OUT/r02s_pytest_gpu.log
tail -6 $OUT/r02s_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/r02s_smoke.log 2>&1; echo "smoke exit $?" >> $OUT/r02s_smoke.log; tail -3 $OUT/r02s_smoke.log
timeout 600 python bench.py --steps 5 --warmup 3 > $OUT/r02s_bench.json 2> $OUT/r02s_bench.err; echo "bench exit $?"; tail -c 2500 $OUT/r02s_bench.json; tail -3 $OUT/r02s_bench.err
for w in config1_massive config5_mnu_w0wa single_cosmology; do
  timeout 600 python bench.py --steps 3 --warmup 3 --workload $w --no-latency > $OUT/r02s_bench_$w.json 2> $OUT/r02s_bench_$w.err; echo "bench $w exit $?"
  python -c "
import json; d=json.load(open('$OUT/r02s_bench_$w.json')); print('$w', d['value'], d['e2e']['value'], d['stage_ms_per_step'], d.get('cpu_baseline',{}).get('value'))" 2>&1 | tail -1
  tail -2 $OUT/r02s_bench_$w.err
done
timeout 600 python bench.py --impl reference --steps 3 --warmup 3 > $OUT/r02s_bench_ref.json 2>> $OUT/r02s_bench.err; tail -c 700 $OUT/r02s_bench_ref.json
