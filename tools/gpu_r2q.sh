#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_scene_gpu.py -m gpu -q -p no:cacheprovider 2>&1 | tail -n 4
timeout 900 python bench.py --no-cpu-baseline > gpurun_out/r2q_bench.json 2> gpurun_out/r2q_bench.err; tail -n 3 gpurun_out/r2q_bench.err
python - <<'PY'
import json
b=json.loads([l for l in open('gpurun_out/r2q_bench.json') if l.startswith('{')][0])
print({k:b[k] for k in ('value','ms_per_step','clocks')})
for k in ('cfg3','cfg4'):
    v=b.get(k)
    if isinstance(v,dict): v={a:c for a,c in v.items() if a not in ('kernels','workload')}
    print(k, json.dumps(v))
PY
