#!/bin/bash
set -u
mkdir -p gpurun_out
for M in 4110 8220 12330 16440; do
  for shape in "1152 384 0" "1536 384 1" "384 384 0"; do
    for lib in libcrossscore_sm100a.so libcrossscore_sm100a_ps.so; do
      XS_LIB_PATH=$PWD/crossscore_b200/$lib timeout 100 python tools/prof_gemm.py $M $shape 2>&1 | tail -n 1 | sed "s/^PAIR=1/$lib/"
    done
  done
done 2>&1 | tee gpurun_out/r2aq_pair_small.txt
