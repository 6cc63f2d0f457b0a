#!/bin/bash
# full GPU check: parity tests, smoke, bench, ncu launch list (run under gpurun)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/pytest_gpu.log
cat gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; cat gpurun_out/smoke.log | tail -3
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.log 2>&1; tail -2 gpurun_out/bench.log
timeout 600 python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/bench_ref.log 2>&1; tail -1 gpurun_out/bench_ref.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 3 --warmup 3 --no-batched --quick > gpurun_out/bench_ncu.log 2>&1
tail -5 gpurun_out/launches.csv
