#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
timeout 900 compute-sanitizer --tool memcheck python scripts/sanitize.py --round2-only > $OUT/r02zzi_memcheck.log 2>&1; tail -2 $OUT/r02zzi_memcheck.log
timeout 1500 compute-sanitizer --tool racecheck python scripts/sanitize.py --round2-only > $OUT/r02zzi_racecheck.log 2>&1; tail -2 $OUT/r02zzi_racecheck.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > $OUT/r02zzi_bench_2gpu.json 2> $OUT/r02zzi_bench_2gpu.err; python -c "
import json; d=json.load(open('$OUT/r02zzi_bench_2gpu.json')); print(d['n_gpus'], d['value'], d['e2e']['value'], d['ms_per_step'])"; tail -2 $OUT/r02zzi_bench_2gpu.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 1 --warmup 1 | tail -1 | cut -c1-300
