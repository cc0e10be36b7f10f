#!/bin/bash
# round 2, final state: GPU suite, smoke, bench lines of every workload, reference arm, launch list of the plain bench command
OUT=gpurun_out; mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -q > $OUT/r02zz_pytest_gpu.log 2>&1; echo "pytest exit $?" >> $OUT/r02zz_pytest_gpu.log
tail -4 $OUT/r02zz_pytest_gpu.log | cut -c1-300
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/r02zz_smoke.log 2>&1; echo "smoke exit $?" >> $OUT/r02zz_smoke.log; tail -3 $OUT/r02zz_smoke.log
timeout 900 python bench.py > $OUT/r02zz_bench.json 2> $OUT/r02zz_bench.err; echo "bench exit $?"; python -c "
import json; d=json.load(open('$OUT/r02zz_bench.json')); print(d['value'], d['e2e'], d['ms_per_step'], d['stage_ms_per_step'], d.get('latency_ms'), d.get('cpu_baseline'), d['roofline']['frac'], d['gpu_launches'], d['clocks'])"; tail -3 $OUT/r02zz_bench.err
timeout 900 python bench.py --batch 64 --no-latency --no-cpu-baseline > $OUT/r02zz_bench_b64.json 2>> $OUT/r02zz_bench.err; python -c "
import json; d=json.load(open('$OUT/r02zz_bench_b64.json')); print('batch 64', d['value'], d['e2e']['value'], d['ms_per_step'], d['stage_ms_per_step'])"
for w in config1_massive config5_mnu_w0wa single_cosmology; do
  timeout 900 python bench.py --steps 3 --warmup 3 --workload $w --no-latency > $OUT/r02zz_bench_$w.json 2> $OUT/r02zz_bench_$w.err; echo "bench $w exit $?"
  python -c "
import json; d=json.load(open('$OUT/r02zz_bench_$w.json')); print('$w', d['value'], d['e2e']['value'], d['stage_ms_per_step'], d.get('cpu_baseline',{}).get('value'), d['roofline']['frac'])" 2>&1 | tail -1
done
timeout 900 python bench.py --impl reference --steps 3 --warmup 3 > $OUT/r02zz_bench_ref.json 2>> $OUT/r02zz_bench.err; tail -c 500 $OUT/r02zz_bench_ref.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/r02zz_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-latency > $OUT/r02zz_ncu_bench.log 2>&1
python - <<'PY'
import csv, collections
rows = list(csv.reader(open('gpurun_out/r02zz_launches.csv')))
hdr = next(r for r in rows if 'Kernel Name' in r)
k, v = hdr.index('Kernel Name'), hdr.index('Metric Value')
agg = collections.Counter(); cnt = collections.Counter()
for r in rows[rows.index(hdr) + 1:]:
    if len(r) > v:
        name = r[k].split('(')[0][:60]
        agg[name] += float(r[v].replace(',', '')); cnt[name] += 1
tot = sum(agg.values())
for n, t in agg.most_common():
    print('%6.2f %%  %8.2f ms  %4d launches  %s' % (100 * t / tot, t / 1e6, cnt[n], n))
PY
