#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -p no:cacheprovider -k "flash_attn or lse or merge" 2>&1 | tail -n 5
for shape in dino192 dec dsa cfg5; do
  SCALE1=1 timeout 120 python tools/prof_attn.py $shape 2>&1 | tail -n 1
done 2>&1 | tee gpurun_out/r2f_attn.txt
XS_LIB_PATH=$PWD/crossscore_b200/libcrossscore_sm100a_prof.so SCALE1=1 timeout 120 python tools/prof_attn.py dino192 2>&1 | tail -n 2 | tee -a gpurun_out/r2f_attn.txt
SCALE1=1 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"attn_tc_kernel" -s 3 -c 1 -o gpurun_out/r2f_prof_attn python tools/prof_attn.py dino192 > gpurun_out/r2f_ncu_attn.log 2>&1
ls -la gpurun_out/r2f_prof_attn.ncu-rep
timeout 900 python -m pytest tests/test_model_gpu.py tests/test_fullsize_gpu.py -m gpu -q -p no:cacheprovider 2>&1 | tail -n 5
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err; python - <<'PY'
import json
b=json.load(open('gpurun_out/r2f_bench.json'))
print({k:b[k] for k in ('value','ms_per_step','roofline','clocks')})
print({k:v for k,v in b['cfg5'].items() if k!='kernels'})
PY
tail -3 gpurun_out/r2f_bench.err
