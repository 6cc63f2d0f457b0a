#!/bin/bash
# round-2 checkpoint: parity tests, bench, source-level ncu capture of the headline kernel
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/pytest_gpu.log
cat gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.log 2>&1; tail -1 gpurun_out/bench.log | cut -c1-1500
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_trsv_systolic -s 4 -c 1 -f -o gpurun_out/r2b_sys python scripts/dev_one.py p3d128 3 > gpurun_out/r2b_sys.log 2>&1
tail -2 gpurun_out/r2b_sys.log
