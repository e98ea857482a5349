#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_fullsize_gpu.py tests/test_predict_gpu.py -m gpu -x -q -p no:cacheprovider 2>&1 | tail -n 5
timeout 600 python bench.py --no-cpu-baseline --no-extra > gpurun_out/r2af_bench.json 2> gpurun_out/r2af_bench.err; tail -n 2 gpurun_out/r2af_bench.err
python - <<'PY'
import json
b=json.loads([l for l in open('gpurun_out/r2af_bench.json') if l.startswith('{')][0])
print({k:b[k] for k in ('value','ms_per_step','clocks','e2e','pipeline')})
PY
