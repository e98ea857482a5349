#!/bin/bash
set -u
mkdir -p gpurun_out
LAYOUT=1 timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"attn_pair_kernel" -s 3 -c 1 -o gpurun_out/${TAG}_prof_attn_pair python tools/prof_attn.py ${SHAPE:-dino192} > gpurun_out/${TAG}_ncu_attn.log 2>&1
tail -n 2 gpurun_out/${TAG}_ncu_attn.log
