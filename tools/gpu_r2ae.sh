#!/bin/bash
# folded LayerNorm: unit tests, plan tests, full suite, bench A/B (XS_FOLD_LN=0/1 in the same run)
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -p no:cacheprovider -k "residual_stats or ln_folded" 2>&1 | tail -n 15
timeout 900 python -m pytest tests/test_fullsize_gpu.py -m gpu -x -q -p no:cacheprovider -k "folded or fused_layernorm" 2>&1 | tail -n 15
timeout 1800 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -n 15 | tee gpurun_out/r2ae_pytest.txt
for rep in 1 2; do
  for f in 0 1; do
    XS_FOLD_LN=$f timeout 600 python bench.py --no-cpu-baseline --no-extra > gpurun_out/r2ae_bench_fold$f.json 2> gpurun_out/r2ae_bench.err
    python - <<PY
import json
b=json.loads([l for l in open('gpurun_out/r2ae_bench_fold$f.json') if l.startswith('{')][0])
print('fold=$f', b['value'], b['ms_per_step'], b['clocks']['sm_mhz'], {t:(v['ms']) for t,v in b['kernels'].items() if 'dino' in t or 'layernorm' in t})
PY
  done
done 2>&1 | tee gpurun_out/r2ae_fold_ab.txt
