# bench.py at N = 1, 2, 4, 8 on one 8-GPU box, the way the driver's scaling run launches it
set -u
mkdir -p gpurun_out
timeout 300 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline 2> gpurun_out/scale_1gpu.err | grep '^{' > gpurun_out/scale_1gpu.json; echo "N=1 rc=$?"
for N in 2 4 8; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2956$N bench.py --gpus $N --steps 20 --warmup 5 2> gpurun_out/scale_${N}gpu.err | grep '^{' > gpurun_out/scale_${N}gpu.json; echo "N=$N rc=$?"
done
python - <<'PY'
import json
base = None
for n in (1, 2, 4, 8):
    try:
        d = json.loads(open(f"gpurun_out/scale_{n}gpu.json").read())
        base = base or d["value"]
        print(n, "value %.4g" % d["value"], "eff %.3f" % (d["value"] / (n * base)), "us/step %.2f" % (d["ms_per_step"] * 1e3),
              "e2e %.4g" % d["e2e"]["value"], "steady %.4g" % d.get("steady_state", {}).get("value", 0))
    except Exception as e:
        print(n, "ERR", e)
PY
