#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -p no:cacheprovider -k "gemm" 2>&1 | tail -n 4
timeout 900 python -m pytest tests/test_model_gpu.py tests/test_fullsize_gpu.py -m gpu -x -q -p no:cacheprovider 2>&1 | tail -n 4
for rep in 1 2; do
  for f in 0 1; do
    XS_TF32_PAIRS=$f timeout 500 python bench.py --no-cpu-baseline --no-extra 2>/dev/null | python -c "
import json,sys
b=json.loads([l for l in sys.stdin if l.startswith('{')][0])
print('tf32_pairs=$f', round(b['value'],1), round(b['ms_per_step'],3), b['clocks']['sm_mhz'], {k:b['kernels'][k]['ms'] for k in ('patch_embed','gemm_dec','head_jigsaw')})"
  done
done 2>&1 | tee gpurun_out/r2an_tf32_pairs_ab.txt
