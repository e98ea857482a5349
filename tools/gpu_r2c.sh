#!/bin/bash
set -u
mkdir -p gpurun_out
for shape in dino192 dec; do
  for lib in prof prof63; do
    XS_LIB_PATH=$PWD/crossscore_b200/libcrossscore_sm100a_$lib.so SCALE1=1 timeout 120 python tools/prof_attn.py $shape 2>&1 | tail -n 3
  done
done 2>&1 | tee gpurun_out/r2c_attn_prof.txt
timeout 600 python -m pytest tests/test_model_gpu.py -m gpu -q -p no:cacheprovider -x 2>&1 | tail -n 5
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2c_bench.json 2> gpurun_out/r2c_bench.err; tail -c 3000 gpurun_out/r2c_bench.json; tail -5 gpurun_out/r2c_bench.err
