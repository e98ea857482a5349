#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -p no:cacheprovider -k "flash_attn" 2>&1 | tail -n 12
for rep in 1 2; do
  for shape in dino192 dec dsa cfg5; do
    for l in 0 1; do LAYOUT=$l timeout 120 python tools/prof_attn.py $shape 2>&1 | tail -n 1; done
  done
done 2>&1 | tee gpurun_out/r2t_attn_layouts.txt
