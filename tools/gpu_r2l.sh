#!/bin/bash
# full GPU suite + the default bench command + its ncu launch list (shares per kernel) for profiles/
set -u
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -n 8 | tee gpurun_out/r2l_pytest.txt
timeout 900 python bench.py > gpurun_out/r2l_bench.json 2> gpurun_out/r2l_bench.err; tail -n 3 gpurun_out/r2l_bench.err
python - <<'PY'
import json
b=json.loads([l for l in open('gpurun_out/r2l_bench.json') if l.startswith('{')][0])
print({k:b[k] for k in ('value','ms_per_step','roofline','clocks','e2e','cpu_baseline')})
for k in ('cfg3','cfg4','cfg5','fp32_mode','torch_gpu'):
    v=b.get(k); 
    if isinstance(v,dict): v={a:c for a,c in v.items() if a!='kernels'}
    print(k, json.dumps(v))
print({t:(v['ms'], v.get('tflops'), v.get('gbs')) for t,v in b['kernels'].items()})
PY
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2l_bench_reference.json 2>/dev/null; cat gpurun_out/r2l_bench_reference.json | cut -c1-600
ncu --metrics gpu__time_duration.sum --clock-control none -s 130 -c 240 --csv --log-file gpurun_out/r2l_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/r2l_ncu_launches.log 2>&1
echo "launch list: $(wc -l < gpurun_out/r2l_launches.csv) lines"
