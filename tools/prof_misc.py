"""ncu driver for the kernels beside the step / flow-field pair: LiDAR scan, learner input, replay / feed push, host export."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import flow_field_based_motion_planner_b200 as ffmp
N = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
dev = torch.device("cuda:0")
env = ffmp.FFMPVectorEnv(N, grid=128, window=100, seed=1234)
env.reset()
acts = torch.randint(0, 28, (40, N), device=dev)
env.rollout(acts)
env.join()
torch.cuda.synchronize()
ring = ffmp.ReplayRing(env, 2)
host_a = torch.randint(0, 28, (N,)).pin_memory()
out = torch.empty((N, 2, 100, 100), dtype=torch.bfloat16, device=dev)
for _ in range(3):
    env.scan(360, 3.5)
    env.learner_input(torch.bfloat16, out=out)
    ring.push(acts[0])
    env.step_host(host_a)
torch.cuda.synchronize()
print("done")
