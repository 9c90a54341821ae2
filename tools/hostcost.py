"""Host issue cost per tick vs device time per tick (step-only workload, no episode ends)."""
import os, sys, time, json
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import flow_field_based_motion_planner_b200 as ffmp

dev = torch.device("cuda:0")
res = {}
for N in (4096, 256):
    env = ffmp.FFMPVectorEnv(N, grid=128, window=100, p_occ=0.0, max_steps=10 ** 9, seed=3)
    env.reset()
    still = torch.full((500, N), 3, dtype=torch.int64, device=dev)
    env.rollout(still); torch.cuda.synchronize()
    t0 = time.perf_counter(); env.rollout(still); t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
    res[f"N{N}_host_issue_us_per_tick"] = (t1 - t0) / 500 * 1e6
    res[f"N{N}_total_us_per_tick"] = (t2 - t0) / 500 * 1e6
    env.close()
print(json.dumps(res))
