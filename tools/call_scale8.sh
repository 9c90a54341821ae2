set -u
mkdir -p gpurun_out
for N in 8; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29560 bench.py --gpus $N --steps 20 --warmup 5 2> gpurun_out/bench_${N}gpu.err | grep '^{' > gpurun_out/bench_${N}gpu.json; echo "bench $N rc=$?"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus $N --config 4 --steps 20 --warmup 5 --no-cpu-baseline 2> gpurun_out/bench_c4_${N}gpu.err | grep '^{' > gpurun_out/bench_c4_${N}gpu.json; echo "c4 $N rc=$?"
done
bash tools/call_loop.sh 8
python - <<'PY'
import json
for f in ("bench_8gpu.json","bench_c4_8gpu.json"):
    try:
        d=json.loads(open("gpurun_out/"+f).read()); print(f, d["value"], d["ms_per_step"], d["e2e"]["value"], d.get("steady_state",{}).get("value"))
    except Exception as e: print(f, "ERR", e)
PY
