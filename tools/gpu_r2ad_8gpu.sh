#!/bin/bash
# 8 GPUs, pair attention layout: bench.py --gpus 8 (headline + cfg3 / cfg4 blocks with parity)
set -u
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r2as_bench_n8.json 2> gpurun_out/r2as_bench_n8.err
tail -n 3 gpurun_out/r2as_bench_n8.err
grep "^{" gpurun_out/r2as_bench_n8.json | python -c "
import json,sys
b=json.loads(sys.stdin.read())
print({k:b[k] for k in ('value','n_gpus','ms_per_step','e2e','clocks')})
print(json.dumps(b.get('cfg3'))); print(json.dumps(b.get('cfg4')))"
