set -u
mkdir -p gpurun_out
timeout 200 python tools/qnet_bench.py 256 > gpurun_out/plain_qnet.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:qnet -c 60 --csv --log-file gpurun_out/launches_qnet.csv python tools/qnet_bench.py 256 > gpurun_out/ncu_qnet.log 2>&1; echo "ncu rc=$?"
cat gpurun_out/plain_qnet.log | tail -2
