#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -p no:cacheprovider -k "flash_attn" 2>&1 | tail -n 5
for shape in dino192 dec dsa; do
  SCALE1=1 timeout 120 python tools/prof_attn.py $shape 2>&1 | tail -n 1
  XS_LIB_PATH=$PWD/crossscore_b200/libcrossscore_sm100a_prof.so SCALE1=1 timeout 120 python tools/prof_attn.py $shape 2>&1 | tail -n 2
done 2>&1 | tee gpurun_out/r2e_attn_prof.txt
timeout 900 python -m pytest tests/test_model_gpu.py -m gpu -q -p no:cacheprovider 2>&1 | tail -n 5
timeout 300 python tools/bench_graph.py 2>&1 | tail -n 1
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2e_bench.json 2> gpurun_out/r2e_bench.err; python - <<'PY'
import json
b=json.load(open('gpurun_out/r2e_bench.json'))
print({k:b[k] for k in ('value','ms_per_step','roofline','kernel_time_share_of_step')})
print(b.get('torch_gpu')); print(b.get('fp32_mode')); print({k:v for k,v in b['cfg5'].items() if k!='kernels'})
PY
tail -3 gpurun_out/r2e_bench.err
