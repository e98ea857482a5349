#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -n 5
timeout 900 python bench.py --no-cpu-baseline > gpurun_out/r2ap_bench.json 2>/dev/null
python - <<'PY'
import json
b=json.loads([l for l in open('gpurun_out/r2ap_bench.json') if l.startswith('{')][0])
print({k:b[k] for k in ('value','ms_per_step','clocks')}, b['e2e']['value'])
for k in ('cfg3','cfg4','cfg5','latency_cfg1'):
    v=b.get(k)
    if isinstance(v,dict): v={a:c for a,c in v.items() if a not in ('kernels','workload','roofline','what','note','parity')}
    print(k, json.dumps(v))
PY
