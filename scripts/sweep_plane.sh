#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 500 python -m pytest tests/test_parity_gpu.py -q -x -k "plane or batched" 2>&1 | tail -5
{
python scripts/dev_batched.py 2048 16
python scripts/dev_batched.py 4096 16
python scripts/dev_batched.py 4096 16 plane_npc=8
python scripts/dev_batched.py 512 16
python scripts/dev_one.py p2d256 10
ILUB_SYS_PLANE=2 python scripts/dev_one.py p2d256 10
} > gpurun_out/sweep_plane.log 2>&1
cat gpurun_out/sweep_plane.log
