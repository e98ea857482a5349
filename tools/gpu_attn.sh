#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -p no:cacheprovider -k "flash_attn" 2>&1 | tail -n 3
echo "--- default mask (0x2AAA = 7/16)"
for shape in dino dec; do timeout 120 python tools/prof_attn.py $shape 2>&1 | tail -n 1; done
for m in 0x0000 0x0888 0x5555 0x6DB6; do
  echo "--- poly mask $m"
  for shape in dino dec; do XS_LIB_PATH=$PWD/tools/ubench_bin_lib_poly$m.so timeout 120 python tools/prof_attn.py $shape 2>&1 | tail -n 1; done
done
timeout 900 python -m pytest tests/test_model_gpu.py -m gpu -x -q -p no:cacheprovider 2>&1 | tail -n 3
