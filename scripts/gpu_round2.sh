#!/bin/bash
# round-2 evidence run (under gpurun, 1 GPU): parity tests, smoke, bench + reference arm, ncu launch list, all configs
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${TAG:-r2m}
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/${T}_pytest_gpu.log; cat gpurun_out/${T}_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${T}_smoke.log 2>&1; tail -2 gpurun_out/${T}_smoke.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/${T}_bench.log 2>&1; tail -1 gpurun_out/${T}_bench.log | cut -c1-400
timeout 600 python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/${T}_bench_ref.log 2>&1; tail -1 gpurun_out/${T}_bench_ref.log | cut -c1-300
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${T}_launches.csv python bench.py --steps 3 --warmup 3 --no-batched --no-extras --quick > gpurun_out/${T}_bench_ncu.log 2>&1
tail -3 gpurun_out/${T}_launches.csv
