#!/bin/bash
# GPU visit r1d: full -m gpu suite, fused-residual A/B bench, packed-LDTM ubench, attention phase clocks
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "[pytest -m gpu] exit $? : $(tail -n 1 gpurun_out/pytest_gpu.log)"
grep -E "^(FAILED|ERROR)|xs:|Error" gpurun_out/pytest_gpu.log | head -20
timeout 120 ./tools/ubench_bin_ldtm > gpurun_out/ubench_ldtm.txt 2>&1; grep -E "pack16|x32 x2 \+ wait|ld \+ 64" gpurun_out/ubench_ldtm.txt
for mode in 1 0; do
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
for shape in dino dec; do
  timeout 120 python tools/prof_attn.py $shape 2>&1 | tail -n 1
  XS_ATTN_PROF=1 timeout 120 python tools/prof_attn.py $shape > gpurun_out/attn_phase_$shape.log 2>&1
  grep -A4 "attn prof" gpurun_out/attn_phase_$shape.log | tail -n 5
done
