#!/bin/bash
# validation with the pair attention layout as default: full GPU suite, smoke(), default bench (+ extras), reference arm, launch list
set -u
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -n 6 | tee gpurun_out/r2ab_pytest.txt
timeout 600 python __graft_entry__.py smoke 2>&1 | tail -n 5
timeout 900 python bench.py > gpurun_out/r2ab_bench.json 2> gpurun_out/r2ab_bench.err; tail -n 3 gpurun_out/r2ab_bench.err
python - <<'PY'
import json
b=json.loads([l for l in open('gpurun_out/r2ab_bench.json') if l.startswith('{')][0])
print({k:b[k] for k in ('value','ms_per_step','roofline','clocks','e2e','gpu_launches','cpu_baseline')})
for k in ('cfg3','cfg4','cfg5','fp32_mode','torch_gpu'):
    v=b.get(k)
    if isinstance(v,dict): v={a:c for a,c in v.items() if a not in ('kernels','workload')}
    print(k, json.dumps(v))
print({t:(v['ms'], v.get('tflops'), v.get('gbs')) for t,v in b['kernels'].items()})
PY
ncu --metrics gpu__time_duration.sum --clock-control none -s 130 -c 240 --csv --log-file gpurun_out/r2ab_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/r2ab_ncu_launches.log 2>&1
echo "launch list: $(wc -l < gpurun_out/r2ab_launches.csv) lines"
