set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "rollout or host_step" > gpurun_out/pytest_graph.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_graph.log
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_graph.json 2> gpurun_out/bench_graph.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_graph.err
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline --no-graph > gpurun_out/bench_nograph.json 2> gpurun_out/bench_nograph.err; echo "bench rc=$?"
python - <<'PY'
import json
for f in ("bench_graph.json", "bench_nograph.json"):
    d = json.loads(open("gpurun_out/" + f).read())
    print(f, "value %.4g" % d["value"], "us/step %.2f" % (d["ms_per_step"] * 1e3), "e2e %.4g" % d["e2e"]["value"], "steady %.4g us %.2f" % (d["steady_state"]["value"], d["steady_state"]["ms_per_step"] * 1e3), "launches", d["gpu_launches"], "roof", round(d["roofline"]["frac"], 3))
PY
