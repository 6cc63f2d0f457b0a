#!/bin/bash
# usage: ncu_one.sh <problem> <tag>   -> gpurun_out/<tag>.ncu-rep (one launch of the systolic kernel, full set)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python scripts/dev_one.py $1 5
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_trsv_systolic -s 4 -c 1 -f -o gpurun_out/$2 python scripts/dev_one.py $1 3 > gpurun_out/$2.log 2>&1
tail -3 gpurun_out/$2.log
