#!/bin/bash
# Closing pass after the last bench.py changes: full GPU suite, smoke, bench lines (both arms, configs 2 / 4), launch list.
set -u
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"
bash tools/call_bench.sh
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches_bench.csv \
    python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1; echo "ncu launches rc=$?"
