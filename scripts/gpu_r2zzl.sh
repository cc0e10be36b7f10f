#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
for rep in 1 2; do for lib in "" variants/lib_gen_nofuse.so; do echo "lib=$lib"; CAMB_B200_LIB=$lib python scripts/evolve_only.py 64 config1_massive 2>&1 | tail -1; CAMB_B200_LIB=$lib python scripts/evolve_only.py 64 config5_mnu_w0wa 2>&1 | tail -1; done; done > $OUT/r02zzl_general_fuse_ab.log 2>&1; cat $OUT/r02zzl_general_fuse_ab.log
