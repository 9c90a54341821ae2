#!/bin/bash
# One GPU-box pass: parity tests, bench (both arms), micro-benchmarks, ncu launch list and full captures.
# Outputs land in gpurun_out/ (scratch); summaries are copied into profiles/ by hand.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
timeout 300 python bench.py --impl reference --steps 100 --warmup 5 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"
timeout 300 python tools/kbench.py > gpurun_out/kbench.json 2> gpurun_out/kbench.err; echo "kbench rc=$?"
timeout 400 python tools/sweep.py > gpurun_out/sweep.txt 2>&1; echo "sweep rc=$?"
timeout 200 python tools/e2e_ab.py 4096 3000 > gpurun_out/e2e_ab.txt 2>&1; echo "e2e_ab rc=$?"
FFMP_HOST_IO_STATS=1 timeout 100 python tools/e2e_ab.py child 4096 3000 > gpurun_out/e2e_stats.txt 2>&1
timeout 120 python tools/feed_bench.py --steps 100 > gpurun_out/feed_bench_1gpu.txt 2>&1; echo "feed rc=$?"
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/plain_bench.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_bench.csv \
    python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1; echo "ncu launches rc=$?"
timeout 120 python tools/prof_step.py 4096 12 > gpurun_out/plain_step.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tick_tma -s 4 -c 2 -o gpurun_out/prof_tick_c -f \
    python tools/prof_step.py 4096 12 > gpurun_out/ncu_tick.log 2>&1; echo "ncu tick rc=$?"
timeout 120 python tools/prof_flow.py 4096 > gpurun_out/plain_flow.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:flow_field_warp -s 1 -c 1 -o gpurun_out/prof_flow_c -f \
    python tools/prof_flow.py 4096 > gpurun_out/ncu_flow.log 2>&1; echo "ncu flow rc=$?"
tail -3 gpurun_out/pytest_gpu.log; cat gpurun_out/bench.json; cat gpurun_out/kbench.json
