#!/bin/bash
# round 2, step b: what binds the attention kernel?  TMEM microbenchmark + timing experiments, then the full GPU suite
set -u
mkdir -p gpurun_out
./tools/ubench_bin_ldtm2 2>&1 | tee gpurun_out/r2b_ubench_ldtm2.txt
for shape in dino192 dec; do
  SCALE1=1 timeout 120 python tools/prof_attn.py $shape 2>&1 | tail -n 1
  for lib in crossscore_b200/libcrossscore_sm100a_d*.so; do
    XS_LIB_PATH=$PWD/$lib SCALE1=1 timeout 120 python tools/prof_attn.py $shape 2>&1 | tail -n 1
  done
done 2>&1 | tee gpurun_out/r2b_attn_dbg.txt
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -n 25
