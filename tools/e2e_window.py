"""Where a 20-step host-buffer window (bench.py's e2e at --steps 20) spends its time: per-step wall clock, the join, the sync."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import flow_field_based_motion_planner_b200 as ffmp
N = 4096
env = ffmp.FFMPVectorEnv(N, grid=128, window=100, seed=1234)
env.reset()
dev = torch.device("cuda:0")
acts = torch.randint(0, 28, (210, N), device=dev)
env.rollout(acts, graph=True); env.rollout(acts, graph=True); env.join(); torch.cuda.synchronize()
host_actions = [torch.randint(0, 28, (N,), dtype=torch.int64).pin_memory() for _ in range(16)]
for i in range(20):
    env.step_host(host_actions[i % 16])
torch.cuda.synchronize()
for rep in range(4):
    ts = [time.perf_counter()]
    for i in range(20):
        env.step_host(host_actions[i % 16])
        ts.append(time.perf_counter())
    env.join(); tj = time.perf_counter()
    torch.cuda.synchronize(); te = time.perf_counter()
    d = [round((b - a) * 1e6, 1) for a, b in zip(ts, ts[1:])]
    print(json.dumps({"steps_us": d, "join_us": round((tj - ts[-1]) * 1e6, 1), "sync_us": round((te - tj) * 1e6, 1),
                      "total_us_per_step": round((te - ts[0]) * 1e6 / 20, 2)}))
    if rep == 1:
        time.sleep(0.5)
