#!/bin/bash
# two GPUs, pair attention layout: the NCCL / peer-memory tests the 1-GPU box skips, then bench.py --gpus 2
set -u
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -n 8 | tee gpurun_out/r2ar_pytest_2gpu.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2ar_bench_n2.json 2> gpurun_out/r2ar_bench_n2.err
tail -n 3 gpurun_out/r2ar_bench_n2.err
grep "^{" gpurun_out/r2ar_bench_n2.json | python -c "
import json,sys
b=json.loads(sys.stdin.read())
print({k:b[k] for k in ('value','n_gpus','ms_per_step','e2e')})
print(json.dumps(b.get('cfg3'))); print(json.dumps(b.get('cfg4')))"
