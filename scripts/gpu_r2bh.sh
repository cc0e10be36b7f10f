#!/bin/bash
# round 2, call BH: evidence for profiles/: launch list of the plain bench command, full ncu captures of the top kernels at the
# bench batch, compute-sanitizer over the round-2 kernels
OUT=gpurun_out; mkdir -p $OUT
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/r02bh_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-latency > $OUT/r02bh_ncu_bench.log 2>&1
python - <<'PY'
import csv, collections
rows = list(csv.reader(open('gpurun_out/r02bh_launches.csv')))
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
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"cb_evolve_reg" -s 9 -c 3 \
    -o $OUT/r02bh_evolve -f python bench.py --steps 1 --warmup 3 --batch 128 --no-cpu-baseline --no-latency > $OUT/r02bh_ncu_evolve.log 2>&1; tail -1 $OUT/r02bh_ncu_evolve.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"cb_los_kernel" -s 6 -c 1 \
    -o $OUT/r02bh_los -f python bench.py --steps 1 --warmup 3 --batch 128 --no-cpu-baseline --no-latency > $OUT/r02bh_ncu_los.log 2>&1; tail -1 $OUT/r02bh_ncu_los.log
timeout 900 compute-sanitizer --tool memcheck python scripts/sanitize.py --round2-only > $OUT/r02bh_memcheck.log 2>&1; tail -3 $OUT/r02bh_memcheck.log
timeout 1200 compute-sanitizer --tool racecheck python scripts/sanitize.py --round2-only > $OUT/r02bh_racecheck.log 2>&1; tail -3 $OUT/r02bh_racecheck.log
