# bench.py at N = 2 and N = 4 on one 4-GPU box, launched the way the driver's scaling run launches it
set -u
mkdir -p gpurun_out
for N in 2 4; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2957$N bench.py --gpus $N --steps 20 --warmup 5 2> gpurun_out/scale_${N}gpu.err | grep '^{' > gpurun_out/scale_${N}gpu.json; echo "N=$N rc=$?"
done
python - <<'PY'
import json
for n in (2, 4):
    try:
        d = json.loads(open(f"gpurun_out/scale_{n}gpu.json").read())
        print(n, "value %.4g" % d["value"], "us/step %.2f" % (d["ms_per_step"] * 1e3), "e2e %.4g" % d["e2e"]["value"], "steady %.4g" % d.get("steady_state", {}).get("value", 0), d.get("ms_per_rank"))
    except Exception as e:
        print(n, "ERR", e)
PY
