#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 3 --warmup 3 --no-latency --no-cpu-baseline > $OUT/r02zzp_bench_2gpu.json 2> $OUT/r02zzp_bench_2gpu.err
wc -l $OUT/r02zzp_bench_2gpu.json; head -c 120 $OUT/r02zzp_bench_2gpu.json; echo; grep -c "NCCL version" $OUT/r02zzp_bench_2gpu.err
python -c "
import json; d=json.load(open('$OUT/r02zzp_bench_2gpu.json')); print(d['n_gpus'], d['value'], d['e2e']['value'])"
