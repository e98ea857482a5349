#!/bin/bash
# ncu evidence for the bench command (B200_PROFILING.md recipe).  Outputs under gpurun_out/.
set -u
mkdir -p gpurun_out
BENCH="python bench.py --steps 1 --warmup 3 --no-cpu-baseline"
# every launch with its device time (cold-cache, serialised: compare SHARES, not absolutes); skip the first forward
ncu --metrics gpu__time_duration.sum --clock-control none -s 130 -c 240 --csv --log-file gpurun_out/launches.csv $BENCH > gpurun_out/ncu_launches.log 2>&1
echo "launch list: $(wc -l < gpurun_out/launches.csv) lines"
# the top kernels, full set, one launch each (after warm-up launches)
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"attn_tc_kernel<.int.4" -s 30 -c 1 -o gpurun_out/prof_attn $BENCH > gpurun_out/ncu_attn.log 2>&1
# the four DINOv2 GEMMs of one layer (qkv, proj+residual, fc1+GELU, fc2+residual: the ring depths 5/6/8 single them out)
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"gemm_tc_kernel<.int.(192|256), .int.[568]," -s 4 -c 4 -o gpurun_out/prof_gemm $BENCH > gpurun_out/ncu_gemm.log 2>&1
ls -la gpurun_out/*.ncu-rep
# the pre-/post-processing kernels, full set, one launch each
ncu --set full --clock-control none --import-source on -k regex:"resize_aa_norm|score_post|u8_norm" -s 12 -c 3 -o gpurun_out/prof_imgproc python tools/prof_imgproc.py > gpurun_out/ncu_imgproc.log 2>&1
ls -la gpurun_out/prof_imgproc.ncu-rep
