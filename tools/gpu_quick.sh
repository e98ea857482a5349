#!/bin/bash
# quick GPU visit: attention kernel tests, model parity, attention timing (+phase clocks), bench
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -p no:cacheprovider -k "${1:-flash_attn}" > gpurun_out/pytest_quick.log 2>&1
echo "[pytest kernels -k ${1:-flash_attn}] exit $? : $(tail -n 1 gpurun_out/pytest_quick.log)"
grep -E "^(FAILED|ERROR)|xs:|Error" gpurun_out/pytest_quick.log | head -20
timeout 120 ./tools/ubench_bin_ldtm > gpurun_out/ubench_ldtm.txt 2>&1; cat gpurun_out/ubench_ldtm.txt
for shape in dino dec; do
  timeout 120 python tools/prof_attn.py $shape 2>&1 | tail -n 1
  XS_ATTN_PROF=1 timeout 120 python tools/prof_attn.py $shape 2>&1 | grep "attn prof" | tail -n 1
done
timeout 900 python -m pytest tests/test_model_gpu.py -m gpu -x -q -p no:cacheprovider > gpurun_out/pytest_model.log 2>&1
echo "[pytest model] exit $? : $(tail -n 1 gpurun_out/pytest_model.log)"
grep -E "^(FAILED|ERROR)|xs:|Error" gpurun_out/pytest_model.log | head -20
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err
echo "[bench] exit $?"; python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/bench_quick.json"))
    print("maps/s", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "ms/step", round(d["ms_per_step"], 2), d["clocks"])
    for k, v in d["kernels"].items():
        print(f"  {k:16s} {v}")
except Exception as e:
    print("bench parse failed", e); print(open("gpurun_out/bench_quick.err").read()[-2000:])
PY
