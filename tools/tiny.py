import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import flow_field_based_motion_planner_b200 as ffmp
env = ffmp.FFMPVectorEnv(8, grid=128, window=100, seed=0)
env.reset()
torch.cuda.synchronize()
print("reset ok")
env.step(torch.zeros(8, dtype=torch.int64, device="cuda:0"))
torch.cuda.synchronize()
print("step ok")
