set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "large or config4" 2>&1 | tail -12
timeout 300 python tools/sweep.py flow 2>&1 | grep "512" | tee gpurun_out/sweep_wide.txt
timeout 300 python tools/sweep.py configs 2>&1 | grep "config4" | tee -a gpurun_out/sweep_wide.txt
