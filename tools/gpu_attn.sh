#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -p no:cacheprovider -k "gemm" 2>&1 | tail -n 5
for dbg in 128 32 64 96 46 110; do
  XS_ATTN_DBG=$dbg timeout 120 python tools/prof_attn.py dino 2>&1 | grep -E "dbg=" | tail -n 1
done
echo "(dbg bits: 1 PV A from smem, 2 no P stores, 4 no S loads, 8 no MUFU, 32 early P arrive, 64 no K/V TMA, 128 nothing)"
for dbg in 128 32 96; do
  XS_ATTN_DBG=$dbg XS_ATTN_PROF=1 timeout 120 python tools/prof_attn.py dino 2>&1 | grep -E "dbg=|attn prof" | tail -n 2
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
XS_GEMM_PAIR=0 timeout 600 python bench.py --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.load(sys.stdin); print('PAIR=0 maps/s', round(d['value'],1)); [print('  ',k,v) for k,v in d['kernels'].items() if k.startswith('gemm_d')]"
