#!/bin/bash
# One gpurun call: GPU parity tests, a bench line, the ncu launch list of the same bench command and one
# full capture of the dominant kernel.  Usage: gpurun --timeout 1500 -- 'bash scripts/gpu_check.sh [tag]'
TAG=${1:-r01}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/${TAG}_smi.txt 2>&1
nproc > $OUT/${TAG}_nproc.txt; (gfortran --version || echo "no gfortran") >> $OUT/${TAG}_nproc.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q -s > $OUT/${TAG}_pytest_gpu.log 2>&1; echo "pytest exit $?" >> $OUT/${TAG}_pytest_gpu.log
tail -5 $OUT/${TAG}_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke.log 2>&1; echo "smoke exit $?" >> $OUT/${TAG}_smoke.log
tail -3 $OUT/${TAG}_smoke.log
timeout 600 python bench.py --steps 5 --warmup 3 > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench exit $?"
tail -c 3000 $OUT/${TAG}_bench.json; tail -5 $OUT/${TAG}_bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/${TAG}_bench_ref.json 2>> $OUT/${TAG}_bench.err
tail -c 600 $OUT/${TAG}_bench_ref.json
# launch list (cold-cache, serialised: shares only)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 3 --batch 16 --no-cpu-baseline > $OUT/${TAG}_ncu_bench.log 2>&1
# full capture of the two main kernels on a small run (smoke: one cosmology, lmax 600)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"cb_evolve_kernel|cb_los_kernel|cb_lens_kernel" -c 3 \
    -o $OUT/${TAG}_prof -f python bench.py --steps 1 --warmup 3 --batch 8 --no-cpu-baseline > $OUT/${TAG}_ncu_full.log 2>&1
ls -la $OUT
