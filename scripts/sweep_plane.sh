#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 500 python -m pytest tests/test_parity_gpu.py -q -x -k "plane or batched" 2>&1 | tail -5
{
python scripts/dev_batched.py 2048 16
python scripts/dev_batched.py 4096 16
python scripts/dev_batched.py 4096 16 plane_npc=8
python scripts/dev_batched.py 2048 16 plane_r=2 plane_wpc=2
python scripts/dev_batched.py 512 16
python scripts/dev_batched.py 4096 1
} > gpurun_out/sweep_plane.log 2>&1
cat gpurun_out/sweep_plane.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_trsv_plane -s 2 -c 1 -f -o gpurun_out/r1_plane python scripts/dev_batched.py 592 16 reps=1 > gpurun_out/ncu_plane.log 2>&1
tail -2 gpurun_out/ncu_plane.log
