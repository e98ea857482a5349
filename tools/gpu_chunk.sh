#!/bin/bash
# chunk-size sweep of the backbone (L2 residency) with and without the fused residual epilogue
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_model_gpu.py -m gpu -x -q -p no:cacheprovider -k "chunked or golden" > gpurun_out/pytest_chunk.log 2>&1
echo "[pytest chunked] exit $? : $(tail -n 1 gpurun_out/pytest_chunk.log)"
grep -E "^(FAILED|ERROR)|xs:|Error" gpurun_out/pytest_chunk.log | head -20
for fuse in 0 1; do for ch in ${CHUNKS:-0 6 8 12 16 24 32 48 96}; do
  XS_FUSE_RESIDUAL=$fuse XS_CHUNK_IMAGES=$ch timeout 300 python bench.py --no-cpu-baseline --steps 6 --warmup 3 > gpurun_out/bench_f${fuse}_c$ch.json 2> gpurun_out/bench_f${fuse}_c$ch.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/bench_f${fuse}_c$ch.json"))
    k = d["kernels"]
    g = lambda n: round(k[n]["ms"], 2) if n in k else None
    print("fuse=$fuse chunk=$ch maps/s", round(d["value"], 1), "ms/step", round(d["ms_per_step"], 2), "clk", d["clocks"]["sm_mhz"],
          "attn", g("attn_dino"), "ln", g("layernorm"), "qkv", g("gemm_dino_qkv"), "proj", g("gemm_dino_proj"), "fc1", g("gemm_dino_fc1"), "fc2", g("gemm_dino_fc2"), "pe", g("patch_embed"))
except Exception as e:
    print("fuse=$fuse chunk=$ch failed", e); print(open("gpurun_out/bench_f${fuse}_c$ch.err").read()[-1500:])
PY
done; done
