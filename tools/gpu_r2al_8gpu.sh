#!/bin/bash
# 8 GPUs: end-to-end (pinned host fp32 in) with and without binding each rank to its GPU's NUMA node
set -u
mkdir -p gpurun_out
lscpu | grep -i "numa\|socket\|^CPU(s)" | head -8
nvidia-smi topo -m 2>/dev/null | head -12 | cut -c1-160
for bind in 0 1 0 1; do
  XS_NUMA_BIND=$bind timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 2951$bind bench.py --gpus 8 --steps 10 --warmup 3 --no-extra > gpurun_out/r2al_bench_n8_bind$bind.json 2> gpurun_out/r2al.err
  grep "^{" gpurun_out/r2al_bench_n8_bind$bind.json | python -c "
import json,sys
b=json.loads(sys.stdin.read())
print('bind=$bind', round(b['value']), {k:(round(v,2) if isinstance(v,float) else v) for k,v in b['e2e'].items()}, round(b['pipeline']['value']))"
done 2>&1 | tee gpurun_out/r2al_numa_ab.txt
