#!/bin/bash
# round 2, step a: new attention kernel -- parity tests, then timing against the round-1 library and poly-mask variants
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv,noheader
timeout 900 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -p no:cacheprovider -k "flash_attn" 2>&1 | tail -n 15
for shape in dino dino192 dec dsa; do
  XS_LIB_PATH=$PWD/tools/ubench_bin_lib_r1.so SCALE1=0 timeout 120 python tools/prof_attn.py $shape 2>&1 | tail -n 1
  SCALE1=0 timeout 120 python tools/prof_attn.py $shape 2>&1 | tail -n 1
  SCALE1=1 timeout 120 python tools/prof_attn.py $shape 2>&1 | tail -n 1
done
for lib in crossscore_b200/libcrossscore_sm100a_*.so; do
  for shape in dino192 dec; do
    XS_LIB_PATH=$PWD/$lib SCALE1=1 timeout 120 python tools/prof_attn.py $shape 2>&1 | tail -n 1
  done
done
timeout 1200 python -m pytest tests/test_model_gpu.py tests/test_kernels_gpu.py -m gpu -x -q -p no:cacheprovider 2>&1 | tail -n 8
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err; tail -c 1500 gpurun_out/r2a_bench.json
