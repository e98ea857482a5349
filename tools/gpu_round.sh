#!/bin/bash
# One GPU visit: full -m gpu suite (as the driver runs it), bench, attention phase clocks, ncu evidence.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
timeout 1200 python -m pytest tests -m gpu -x -q -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "[pytest -m gpu] exit $? : $(tail -n 1 gpurun_out/pytest_gpu.log)"
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "[bench] exit $?"; cut -c1-600 gpurun_out/bench.json
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2>> gpurun_out/bench.err
echo "[bench ref] exit $?"; cut -c1-300 gpurun_out/bench_ref.json
for shape in dino dec; do
  XS_ATTN_PROF=1 timeout 120 python tools/prof_attn.py $shape > gpurun_out/attn_phase_$shape.log 2>&1
  tail -n 3 gpurun_out/attn_phase_$shape.log
  timeout 120 python tools/prof_attn.py $shape 2>&1 | tail -n 1
done
if [ "${1:-}" = "ncu" ]; then bash tools/gpu_profile.sh; fi
