#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -p no:cacheprovider -k "gemm" 2>&1 | tail -n 8
timeout 900 python -m pytest tests/test_model_gpu.py -m gpu -x -q -p no:cacheprovider 2>&1 | tail -n 5
for rep in 1 2; do
  for f in 1 0; do
    XS_FUSE_LN=$f timeout 600 python bench.py --steps 10 --warmup 3 --no-extra --no-cpu-baseline > gpurun_out/r2k_bench_ln$f.json 2> gpurun_out/r2k_bench_ln$f.err
    python - <<PY
import json
b=json.loads([l for l in open('gpurun_out/r2k_bench_ln$f.json') if l.startswith('{')][0])
k=b['kernels']
print('fuse_ln=$f', round(b['value'],1), 'maps/s', round(b['ms_per_step'],3), 'ms', {t:(k[t]['ms']) for t in k if 'proj' in t or 'fc2' in t or t=='layernorm'})
PY
  done
done
