#!/bin/bash
# round 2, call G: Tensor-Memory probe, parity of the mode-per-lane kernel, bench
OUT=gpurun_out; mkdir -p $OUT
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -s -k "tensor_memory" > $OUT/r02g_probe.log 2>&1; echo "probe exit $?" >> $OUT/r02g_probe.log
tail -4 $OUT/r02g_probe.log
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -s -k "planck_massless or cl_parity" > $OUT/r02g_parity1.log 2>&1; echo "parity1 exit $?" >> $OUT/r02g_parity1.log
tail -15 $OUT/r02g_parity1.log
timeout 900 python -m pytest tests -m gpu -q > $OUT/r02g_pytest_gpu.log 2>&1; echo "pytest exit $?" >> $OUT/r02g_pytest_gpu.log
tail -12 $OUT/r02g_pytest_gpu.log
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $OUT/r02g_bench.json 2> $OUT/r02g_bench.err; tail -c 1500 $OUT/r02g_bench.json; tail -3 $OUT/r02g_bench.err
