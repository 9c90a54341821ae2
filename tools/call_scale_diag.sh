set -u
mkdir -p gpurun_out
run() { # name, extra env, args
  env $2 timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $3 bench.py --gpus 8 --warmup 5 --no-extras --no-cpu-baseline $4 2> gpurun_out/diag_$1.err | grep '^{' > gpurun_out/diag_$1.json
  python - "$1" <<'PY'
import json, sys
d = json.loads(open(f"gpurun_out/diag_{sys.argv[1]}.json").read())
print(sys.argv[1], "us/step %.2f" % (d["ms_per_step"] * 1e3), "per rank ms", d["ms_per_rank"])
PY
}
run base "A=1" 29571 "--steps 20"
run nosampler "BENCH_NO_SAMPLER=1" 29572 "--steps 20"
run steps80 "BENCH_NO_SAMPLER=1" 29573 "--steps 80"
run nograph "BENCH_NO_SAMPLER=1" 29574 "--steps 20 --no-graph"
run base2 "A=1" 29575 "--steps 20"
