#!/bin/bash
# dev sweep: option sensitivity of the sync-free solve + one ncu capture
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
B="timeout 120 python scripts/bench_apply.py"
{
$B p3d128 ichol0
$B p3d128 ichol0 near_window=0
$B p3d128 ichol0 warps_per_cta=8
$B p3d128 ichol0 warps_per_cta=32
$B p3d128 ichol0 ctas_per_sm=2 warps_per_cta=16 ring_chunks=128 near_window=64
$B p3d128 ichol0 ctas_per_sm=4 warps_per_cta=8 ring_chunks=64 near_window=32
$B p3d128 ichol0 stages=1
$B p3d128 ichol0 exact=0
$B p3d128 ichol0 solve_algo=1
$B p2d256 ilu0
$B p2d256 ilu0 near_window=0
$B p2d256 ilu0 solve_algo=1
$B p3d64 ichol0
} 2>&1 | tee gpurun_out/sweep1.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_trsv_syncfree -s 6 -c 2 -o gpurun_out/trsv_r1a python scripts/bench_apply.py p3d128 ichol0 reps=3 > gpurun_out/ncu1.log 2>&1
tail -3 gpurun_out/ncu1.log
