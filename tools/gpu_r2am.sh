#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -p no:cacheprovider -k "flash_attn and layout_pair" 2>&1 | tail -n 4
for rep in 1 2; do
  for lib in ${LIBS}; do
    for shape in dino192 dec cfg5; do
      XS_LIB_PATH=$PWD/crossscore_b200/$lib LAYOUT=1 timeout 120 python tools/prof_attn.py $shape 2>&1 | tail -n 1
    done
  done
done 2>&1 | tee gpurun_out/${TAG}_attn_ab.txt
