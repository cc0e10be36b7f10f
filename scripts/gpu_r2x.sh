#!/bin/bash
# round 2, call X: fused-output build against the unfused one, same box: bench + per-kernel launch list
OUT=gpurun_out; mkdir -p $OUT
bash scripts/variants_bench.sh r02x_variants
for lib in camb_b200/libcambb200.so variants/libcambb200_unfused.so; do
  tag=$(basename $lib .so)
  CAMB_B200_LIB=$PWD/$lib timeout 600 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio --clock-control none -k regex:"cb_evolve" -c 12 --csv --log-file $OUT/r02x_launches_$tag.csv \
    python bench.py --steps 1 --warmup 1 --batch 64 --no-cpu-baseline --no-latency > $OUT/r02x_ncu_$tag.log 2>&1
  echo "== $tag"; grep cb_evolve $OUT/r02x_launches_$tag.csv | awk -F'","' '{print substr($5,1,60), $(NF-2), $NF}' | tail -9
done
