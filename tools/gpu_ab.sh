#!/bin/bash
# A/B visit: kernel + model tests, then bench with/without the fused residual epilogue
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_model_gpu.py -m gpu -x -q -p no:cacheprovider > gpurun_out/pytest_ab.log 2>&1
echo "[pytest kernels+model] exit $? : $(tail -n 1 gpurun_out/pytest_ab.log)"
grep -E "^(FAILED|ERROR)|xs:|Error" gpurun_out/pytest_ab.log | head -20
for mode in ${MODES:-1 0}; do
  XS_FUSE_RESIDUAL=$mode timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_fuse$mode.json 2> gpurun_out/bench_fuse$mode.err
  echo "[bench fuse=$mode] exit $?"; python - <<PY
import json
try:
    d = json.load(open("gpurun_out/bench_fuse$mode.json"))
    print("maps/s", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "ms/step", round(d["ms_per_step"], 2), d["clocks"])
    for k, v in d["kernels"].items():
        print(f"  {k:16s} {v}")
except Exception as e:
    print("bench parse failed", e); print(open("gpurun_out/bench_fuse$mode.err").read()[-2000:])
PY
done
