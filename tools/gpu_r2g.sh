#!/bin/bash
# in-run A/B of attention variants (box-to-box noise is +-3 %: compare only within one run, interleaved repetitions)
set -u
mkdir -p gpurun_out
for rep in 1 2 3; do
  SCALE1=1 timeout 120 python tools/prof_attn.py dino192 2>&1 | tail -n 1
  for lib in nd nd5 d5 d3 d0; do
    XS_LIB_PATH=$PWD/crossscore_b200/libcrossscore_sm100a_$lib.so SCALE1=1 timeout 120 python tools/prof_attn.py dino192 2>&1 | tail -n 1
  done
done 2>&1 | tee gpurun_out/r2g_attn_ab.txt
for lib in "" _nd _d5; do
  XS_LIB_PATH=$PWD/crossscore_b200/libcrossscore_sm100a$lib.so SCALE1=1 timeout 120 python tools/prof_attn.py dec 2>&1 | tail -n 1
done 2>&1 | tee -a gpurun_out/r2g_attn_ab.txt
