# bench.py at N = 1 and N = 8 on one 8-GPU box, launched the way the driver's scaling run launches it
set -u
mkdir -p gpurun_out
timeout 300 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline 2> gpurun_out/scale_1gpu.err | grep '^{' > gpurun_out/scale_1gpu.json; echo "N=1 rc=$?"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29568 bench.py --gpus 8 --steps 20 --warmup 5 2> gpurun_out/scale_8gpu.err | grep '^{' > gpurun_out/scale_8gpu.json; echo "N=8 rc=$?"
python - <<'PY'
import json
base = None
for n in (1, 8):
    try:
        d = json.loads(open(f"gpurun_out/scale_{n}gpu.json").read())
        base = base or d["value"]
        print(n, "value %.4g" % d["value"], "eff %.3f" % (d["value"] / (n * base)), "us/step %.2f" % (d["ms_per_step"] * 1e3),
              "e2e %.4g" % d["e2e"]["value"], "steady %.4g" % d.get("steady_state", {}).get("value", 0), d.get("ms_per_rank"))
    except Exception as e:
        print(n, "ERR", e)
PY
tail -3 gpurun_out/scale_8gpu.err
