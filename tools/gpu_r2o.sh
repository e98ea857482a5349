#!/bin/bash
# final validation of the round: full GPU suite, smoke(), default bench, ncu --set full of the shipped attention kernel
set -u
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -n 6 | tee gpurun_out/r2o_pytest.txt
timeout 600 python __graft_entry__.py smoke 2>&1 | tail -n 5
timeout 900 python bench.py > gpurun_out/r2o_bench.json 2> gpurun_out/r2o_bench.err; tail -n 3 gpurun_out/r2o_bench.err
python - <<'PY'
import json
b=json.loads([l for l in open('gpurun_out/r2o_bench.json') if l.startswith('{')][0])
print({k:b[k] for k in ('value','ms_per_step','roofline','clocks','e2e','gpu_launches')})
for k in ('cfg3','cfg4','cfg5'):
    v=b.get(k)
    if isinstance(v,dict): v={a:c for a,c in v.items() if a not in ('kernels','workload')}
    print(k, json.dumps(v))
print({t:(v['ms'], v.get('tflops'), v.get('gbs')) for t,v in b['kernels'].items()})
PY
SCALE1=1 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"attn_tc_kernel" -s 3 -c 1 -o gpurun_out/r2o_prof_attn python tools/prof_attn.py dino192 > gpurun_out/r2o_ncu_attn.log 2>&1
ls -la gpurun_out/r2o_prof_attn.ncu-rep
