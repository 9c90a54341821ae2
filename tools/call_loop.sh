set -u
mkdir -p gpurun_out
NG=${1:-2}
for feed in push nccl none; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29552 tools/train_loop.py --envs 256 --steps 24 --feed $feed 2>&1 | grep '^{' | tee -a gpurun_out/train_loop_${NG}gpu.txt
done
