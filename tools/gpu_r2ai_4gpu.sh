#!/bin/bash
set -u
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 4 --steps 10 --warmup 3 > gpurun_out/r2ai_bench_n4.json 2> gpurun_out/r2ai_bench_n4.err
tail -n 3 gpurun_out/r2ai_bench_n4.err
grep "^{" gpurun_out/r2ai_bench_n4.json | python -c "
import json,sys
b=json.loads(sys.stdin.read())
print({k:b[k] for k in ('value','n_gpus','ms_per_step','e2e')})
print(json.dumps(b.get('cfg3'))); print(json.dumps(b.get('cfg4')))"
