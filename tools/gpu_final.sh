#!/bin/bash
# what the driver runs at round end (suite, smoke, bench both arms) + the GEMM ncu capture
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "[pytest -m gpu] exit $? : $(tail -n 1 gpurun_out/pytest_gpu.log)"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2>/dev/null; echo "[bench ref] exit $?"
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "[bench] exit $?"
BENCH="python bench.py --steps 1 --warmup 3 --no-cpu-baseline"
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"gemm_tc_kernel<.int.(192|256), .int.[568]," -s 4 -c 4 -o gpurun_out/prof_gemm $BENCH > gpurun_out/ncu_gemm.log 2>&1
ls -la gpurun_out/prof_gemm.ncu-rep
