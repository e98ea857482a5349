#!/bin/bash
# fp16-logit attention: kernel tests, model parity, kernel timing + phase clocks, bench A/B
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -p no:cacheprovider -k "f16 or attn_probs or flash" > gpurun_out/pytest_f16.log 2>&1
echo "[pytest kernels f16] exit $? : $(tail -n 1 gpurun_out/pytest_f16.log)"
grep -E "^(FAILED|ERROR)|xs:|Error|assert|max err" gpurun_out/pytest_f16.log | head -20
for shape in dino dec; do
  F16=0 timeout 120 python tools/prof_attn.py $shape 2>&1 | tail -n 1
  F16=1 timeout 120 python tools/prof_attn.py $shape 2>&1 | tail -n 1
  F16=1 XS_ATTN_PROF=1 timeout 120 python tools/prof_attn.py $shape 2>&1 | grep "attn f16 prof" | tail -n 1
done
timeout 900 python -m pytest tests/test_model_gpu.py tests/test_scene_gpu.py -m gpu -x -q -p no:cacheprovider > gpurun_out/pytest_model.log 2>&1
echo "[pytest model+scene] exit $? : $(tail -n 1 gpurun_out/pytest_model.log)"
grep -E "^(FAILED|ERROR)|xs:|Error|assert" gpurun_out/pytest_model.log | head -20
for mode in 1 0; do
  XS_ATTN_F16=$mode timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_f16_$mode.json 2> gpurun_out/bench_f16_$mode.err
  echo "[bench attn_f16=$mode] exit $?"; python - <<PY
import json
try:
    d = json.load(open("gpurun_out/bench_f16_$mode.json"))
    print("maps/s", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "ms/step", round(d["ms_per_step"], 2), d["clocks"])
    for k, v in d["kernels"].items():
        print(f"  {k:16s} {v}")
except Exception as e:
    print("bench parse failed", e); print(open("gpurun_out/bench_f16_$mode.err").read()[-2000:])
PY
done
