#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
t0=$(date +%s)
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 4 --steps 5 --warmup 3 > $OUT/r02zzm_bench_4gpu.json 2> $OUT/r02zzm_bench_4gpu.err
echo "wall $(( $(date +%s) - t0 )) s"
python -c "
import json; t=open('$OUT/r02zzm_bench_4gpu.json').read(); d=json.loads(t[t.index('{'):]); print(d['n_gpus'], d['value'], d['e2e']['value'], d['ms_per_step'], d['e2e']['step_wall_ms'])"; tail -2 $OUT/r02zzm_bench_4gpu.err
