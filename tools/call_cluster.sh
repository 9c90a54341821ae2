set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "cluster_variant or large_map or large_without" > gpurun_out/pytest_cluster.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/pytest_cluster.log
timeout 600 python tools/cluster_ab.py > gpurun_out/cluster_ab.txt 2>&1; cat gpurun_out/cluster_ab.txt
