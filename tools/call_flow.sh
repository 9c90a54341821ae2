set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "flow_field or scenarios or rollout_1k or interleaved or config1 or config2" > gpurun_out/pytest_flow.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/pytest_flow.log
timeout 300 python tools/flow_ab.py > gpurun_out/flow_ab.txt 2>&1; cat gpurun_out/flow_ab.txt
timeout 120 python tools/prof_flow.py 4096 > gpurun_out/plain_flow.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:flow_field_il -s 1 -c 1 -o gpurun_out/prof_flow_il -f \
    python tools/prof_flow.py 4096 > gpurun_out/ncu_flow.log 2>&1; echo "ncu flow rc=$?"
