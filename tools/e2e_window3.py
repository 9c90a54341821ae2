"""How many host-buffer steps a fresh process needs before the step settles (development probe): consecutive 100-step windows."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import flow_field_based_motion_planner_b200 as ffmp
N = 4096
env = ffmp.FFMPVectorEnv(N, grid=128, window=100, seed=1234, ring=32, slots=16, device="cuda:0")
env.reset()
torch.cuda.synchronize()
time.sleep(1.0)
host_actions = [torch.randint(0, 28, (N,), dtype=torch.int64).pin_memory() for _ in range(16)]
w = []
for k in range(40):
    t0 = time.perf_counter()
    for i in range(100):
        env.step_host(host_actions[i % 16])
    w.append(round((time.perf_counter() - t0) * 1e4, 1))
print(json.dumps({"us_per_step_by_100_step_window": w}))
