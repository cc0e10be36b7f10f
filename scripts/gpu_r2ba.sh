#!/bin/bash
# round 2, call BA: the state of the tree: GPU suite, smoke, bench lines of every workload, reference arm
OUT=gpurun_out; mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -q > $OUT/r02ba_pytest_gpu.log 2>&1; echo "pytest exit $?" >> $OUT/r02ba_pytest_gpu.log
tail -4 $OUT/r02ba_pytest_gpu.log | cut -c1-300
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/r02ba_smoke.log 2>&1; echo "smoke exit $?" >> $OUT/r02ba_smoke.log; tail -3 $OUT/r02ba_smoke.log
timeout 900 python bench.py --steps 5 --warmup 3 > $OUT/r02ba_bench.json 2> $OUT/r02ba_bench.err; echo "bench exit $?"; python -c "
import json; d=json.load(open('$OUT/r02ba_bench.json')); print(d['value'], d['e2e'], d['ms_per_step'], d['stage_ms_per_step'], d.get('latency_ms'), d.get('cpu_baseline'))"; tail -3 $OUT/r02ba_bench.err
for w in config1_massive config5_mnu_w0wa single_cosmology; do
  timeout 600 python bench.py --steps 3 --warmup 3 --workload $w --no-latency > $OUT/r02ba_bench_$w.json 2> $OUT/r02ba_bench_$w.err; echo "bench $w exit $?"
  python -c "
import json; d=json.load(open('$OUT/r02ba_bench_$w.json')); print('$w', d['value'], d['e2e']['value'], d['stage_ms_per_step'], d.get('cpu_baseline',{}).get('value'))" 2>&1 | tail -1
done
timeout 600 python bench.py --impl reference --steps 3 --warmup 3 > $OUT/r02ba_bench_ref.json 2>> $OUT/r02ba_bench.err; tail -c 600 $OUT/r02ba_bench_ref.json
