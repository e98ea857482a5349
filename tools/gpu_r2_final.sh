#!/bin/bash
# final validation of the round: full GPU suite, smoke(), default bench, reference arm
set -u
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -n 6 | tee gpurun_out/r2final_pytest.txt
timeout 600 python __graft_entry__.py smoke 2>&1 | tail -n 5
timeout 900 python bench.py > gpurun_out/r2final_bench.json 2> gpurun_out/r2final_bench.err; tail -n 3 gpurun_out/r2final_bench.err
python - <<'PY'
import json
b=json.loads([l for l in open('gpurun_out/r2final_bench.json') if l.startswith('{')][0])
print({k:b[k] for k in ('value','ms_per_step','roofline','clocks','e2e','gpu_launches')})
print({k:(b[k]['value'] if 'value' in b[k] else None) for k in ('pipeline','cpu_baseline')})
for k in ('cfg3','cfg4','cfg5','fp32_mode','torch_gpu','latency_cfg1'):
    v=b.get(k)
    if isinstance(v,dict): v={a:c for a,c in v.items() if a not in ('kernels','workload','roofline','what','note')}
    print(k, json.dumps(v))
PY
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2final_bench_reference.json 2>/dev/null; cut -c1-500 gpurun_out/r2final_bench_reference.json
