#!/bin/bash
# compute-sanitizer memcheck over the pair-layout attention tests and the new folded-LayerNorm GEMM tests
set -u
mkdir -p gpurun_out
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 3 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -p no:cacheprovider -k "layout_pair and (flash_attn_bf16_tc or f32_out) and not long_kv and not adversarial" 2>&1 | tail -n 8 | tee gpurun_out/r2aj_sanitizer_attn.txt
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 3 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -p no:cacheprovider -k "residual_stats or ln_folded" 2>&1 | tail -n 8 | tee gpurun_out/r2aj_sanitizer_gemm.txt
