#!/bin/bash
# one full-set ncu capture of the dominant kernel of the bench workload (run under gpurun, 1 GPU)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_trsv_systolic -s 12 -c 2 -f \
    -o gpurun_out/r1_trsv_systolic python bench.py --steps 3 --warmup 3 --no-batched --quick > gpurun_out/ncu_top.log 2>&1
tail -2 gpurun_out/ncu_top.log | cut -c1-300
