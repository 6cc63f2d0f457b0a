#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
B="timeout 120 python scripts/bench_apply.py"
{
timeout 600 python scripts/dev_gpu_check.py quick 2>&1 | grep -E "FAIL|Traceback|Error|apply:|rror" | tail -30
$B p3d128 ichol0 check=1
$B p3d128 ichol0 exact=0 check=1
$B p3d128 ichol0 exact=0 fastpath=0
$B p3d128 ichol0 exact=0 level_align=0
$B p3d128 ichol0 exact=0 gate_slack=-1
$B p3d128 ichol0 exact=0 warps_per_cta=8
$B p3d128 ilu0 exact=0
$B p2d256 ilu0 exact=0
$B p2d256 ilu0 exact=0 nparts=4
$B p2d256 ilu0 exact=0 nparts=8 level_align=1
$B p2d256 ilu0 exact=0 nparts=2 level_align=1
$B p3d64 ichol0 exact=0
$B cd3d160 ilu0 exact=0 check=1
} 2>&1 | tee gpurun_out/sweep6.log
python scripts/trace_solve.py p3d128 ichol0 exact=0 2>&1 | tail -42
