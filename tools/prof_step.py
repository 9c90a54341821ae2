"""Tiny driver for ncu: a few steps of the bench workload (and a no-episode-end variant)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import flow_field_based_motion_planner_b200 as ffmp  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 12
dev = torch.device("cuda:0")
env = ffmp.FFMPVectorEnv(N, grid=128, window=100, seed=1234)
env.reset()
acts = torch.randint(0, 28, (steps, N), device=dev)
for t in range(steps):
    env.step(acts[t])
env.join()
torch.cuda.synchronize()
env.close()
print("done")
