#!/bin/bash
set -u
for rep in 1 2; do
  XS_LIB_PATH=$PWD/tools/ubench_bin_lib_base.so timeout 300 python tools/bench_cfg3_batch.py 2>&1 | tail -n 1
  timeout 300 python tools/bench_cfg3_batch.py 2>&1 | tail -n 1
done
