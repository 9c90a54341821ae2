set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "host_step or invalid_actions or rollout_api or warp_kernel_regen" > gpurun_out/pytest_act.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/pytest_act.log
timeout 600 python tools/act_ab.py > gpurun_out/act_ab.txt 2>&1; echo "act_ab rc=$?"
cat gpurun_out/act_ab.txt
