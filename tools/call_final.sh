#!/bin/bash
# Round-2 closing pass on one B200: parity tests, bench (both arms, configs 2 / 4), ncu launch list, ncu --set full captures.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
timeout 300 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"
timeout 300 python bench.py --config 2 --steps 200 --warmup 20 --no-cpu-baseline > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; echo "c2 rc=$?"
timeout 300 python bench.py --config 4 --steps 40 --warmup 5 --no-cpu-baseline > gpurun_out/bench_c4.json 2> gpurun_out/bench_c4.err; echo "c4 rc=$?"
timeout 300 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench default rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_bench.csv \
    python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1; echo "ncu launches rc=$?"
timeout 120 python tools/prof_step.py 4096 12 > gpurun_out/plain_step.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tick_tma -s 6 -c 1 -o gpurun_out/prof_tick -f \
    python tools/prof_step.py 4096 12 > gpurun_out/ncu_tick.log 2>&1; echo "ncu tick rc=$?"
timeout 120 python tools/prof_flow.py 4096 > gpurun_out/plain_flow.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:flow_field_il -s 1 -c 1 -o gpurun_out/prof_flow -f \
    python tools/prof_flow.py 4096 > gpurun_out/ncu_flow.log 2>&1; echo "ncu flow rc=$?"
timeout 120 python tools/prof_quad.py 296 > gpurun_out/plain_quad.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:flow_field_quad -s 4 -c 1 -o gpurun_out/prof_quad -f \
    python tools/prof_quad.py 296 > gpurun_out/ncu_quad.log 2>&1; echo "ncu quad rc=$?"
cat gpurun_out/bench.json | cut -c1-1500
