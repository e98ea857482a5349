#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -p no:cacheprovider -k "gemm" 2>&1 | tail -n 2
M=263040
for pm in 0 1 2; do
  for shape in "1152 384 0" "384 384 0" "1536 384 1" "384 1536 0"; do
    XS_GEMM_PAIR=$pm timeout 60 python tools/prof_gemm.py $M $shape 2>&1 | tail -n 1
  done
done
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
