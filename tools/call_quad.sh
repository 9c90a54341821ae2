set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "rollout or flow_field or single_env or masked_reset or checkpoint" > gpurun_out/pytest_quad.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/pytest_quad.log
timeout 600 python tools/quad_ab.py > gpurun_out/quad_ab.txt 2>&1; echo "quad_ab rc=$?"
cat gpurun_out/quad_ab.txt
