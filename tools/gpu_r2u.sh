#!/bin/bash
# pair layout as the default: full GPU suite, short bench, knob A/B, ncu of the pair kernel
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q -p no:cacheprovider 2>&1 | tail -n 5 | tee gpurun_out/r2u_pytest.txt
timeout 600 python bench.py --no-cpu-baseline --no-extra > gpurun_out/r2u_bench.json 2> gpurun_out/r2u_bench.err; tail -n 2 gpurun_out/r2u_bench.err
python - <<'PY'
import json
b=json.loads([l for l in open('gpurun_out/r2u_bench.json') if l.startswith('{')][0])
print({k:b[k] for k in ('value','ms_per_step','clocks','roofline','e2e')})
PY
for rep in 1 2; do
  for lib in libcrossscore_sm100a.so libcrossscore_sm100a_p4.so libcrossscore_sm100a_p6.so libcrossscore_sm100a_p8.so libcrossscore_sm100a_st3.so libcrossscore_sm100a_st5.so; do
    for shape in dino192 dec; do
      XS_LIB_PATH=$PWD/crossscore_b200/$lib LAYOUT=1 timeout 120 python tools/prof_attn.py $shape 2>&1 | tail -n 1
    done
  done
done 2>&1 | tee gpurun_out/r2u_attn_knobs.txt
LAYOUT=1 timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"attn_pair_kernel" -s 3 -c 1 -o gpurun_out/r2u_prof_attn_pair python tools/prof_attn.py dino192 > gpurun_out/r2u_ncu_attn.log 2>&1
tail -n 2 gpurun_out/r2u_ncu_attn.log
