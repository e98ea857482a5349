#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -p no:cacheprovider -k "flash_attn" 2>&1 | tail -n 3
timeout 600 python -m pytest tests/test_model_gpu.py tests/test_scene_gpu.py -m gpu -x -q -p no:cacheprovider 2>&1 | tail -n 3
for rep in 1 2 3; do
  for lib in tools/ubench_bin_lib_base.so crossscore_b200/libcrossscore_sm100a.so crossscore_b200/libcrossscore_sm100a_p5.so crossscore_b200/libcrossscore_sm100a_p6.so; do
    XS_LIB_PATH=$PWD/$lib SCALE1=1 timeout 120 python tools/prof_attn.py dino192 2>&1 | tail -n 1
  done
done 2>&1 | tee gpurun_out/r2r_attn_ab.txt
for lib in tools/ubench_bin_lib_base.so crossscore_b200/libcrossscore_sm100a.so crossscore_b200/libcrossscore_sm100a_p5.so; do
  XS_LIB_PATH=$PWD/$lib SCALE1=1 timeout 120 python tools/prof_attn.py dec 2>&1 | tail -n 1
done 2>&1 | tee -a gpurun_out/r2r_attn_ab.txt
