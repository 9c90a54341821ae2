"""ncu driver: flow_field_quad_kernel (four warps per grid) on a reset of 74 envs x 4 scenario slots = 296 grids in one launch."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import flow_field_based_motion_planner_b200 as ffmp
N = int(sys.argv[1]) if len(sys.argv) > 1 else 74
env = ffmp.FFMPVectorEnv(N, grid=128, window=100, seed=1234, slots=4)
env.reset()
env.reset()
torch.cuda.synchronize()
print("done", env.error_word())
env.close()
