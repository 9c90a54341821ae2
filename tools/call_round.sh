set -u
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
timeout 300 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"
timeout 300 python bench.py --config 2 --steps 200 --warmup 20 --no-cpu-baseline > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; echo "c2 rc=$?"
timeout 300 python bench.py --config 4 --steps 40 --warmup 5 --no-cpu-baseline > gpurun_out/bench_c4.json 2> gpurun_out/bench_c4.err; echo "c4 rc=$?"
cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err; cat gpurun_out/bench_c2.json | cut -c1-600; cat gpurun_out/bench_c4.json | cut -c1-600; tail -3 gpurun_out/bench_c4.err
