#!/bin/bash
# Run the GPU kernel tests one test-function per process (a device trap poisons the CUDA context, so
# isolation keeps the rest of the suite informative).  Logs go to gpurun_out/.
#   tools/gpu_suite.sh kernels|model|all
set -u
mkdir -p gpurun_out
what=${1:-all}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
run_group () {  # file, -k expression, log name
  timeout 600 python -m pytest "$1" -q -x -m gpu -k "$2" -p no:cacheprovider > "gpurun_out/$3.log" 2>&1
  echo "[$3] exit $? : $(tail -n 1 gpurun_out/$3.log)"
}
if [ "$what" = "kernels" ] || [ "$what" = "all" ]; then
  for t in test_library_and_device test_gemm_f32 test_gemm_bf16_tc test_gemm_tf32_tc test_gemm_bf16_in_f32_out test_gemm_bf16_rejects test_flash_attn_f32 \
           "test_flash_attn_bf16_tc and not f32_out" test_flash_attn_bf16_tc_f32_out test_layernorm_residual test_patch_embed test_embed_and_final_ln test_table_resamplers \
           test_head_score_jigsaw test_attn_probs_one_head; do
    run_group tests/test_kernels_gpu.py "$t" "k_$(echo $t | tr ' ' '_')"
  done
fi
if [ "$what" = "model" ] || [ "$what" = "all" ]; then
  run_group tests/test_model_gpu.py "test_golden_parity and fp32" m_golden_fp32
  run_group tests/test_model_gpu.py "test_golden_parity and bf16" m_golden_bf16
  run_group tests/test_model_gpu.py "test_featmaps or test_batch_vs_oracle or test_errors" m_misc
  run_group tests/test_model_gpu.py "test_headline" m_headline
fi
grep -h -E "^(FAILED|ERROR)|Error|error:|assert|xs:" gpurun_out/*.log | head -60
