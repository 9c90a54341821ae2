"""Observation ring length K (frames [N][K][W][W]): on a ring wrap (every K-1 steps) every env rewrites its older frame as well,
so the step's frame writes are (1 + 1/(K-1)) x 10 KB per env.  Steady step / 20-step window of the bench workload vs K."""
import json, os, subprocess, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if len(sys.argv) > 1 and sys.argv[1] == "child":
    import math
    import torch
    import flow_field_based_motion_planner_b200 as ffmp
    K = int(sys.argv[2])
    dev = torch.device("cuda:0")
    env = ffmp.FFMPVectorEnv(4096, grid=128, window=100, seed=1234, ring=K, slots=16)
    env.reset()
    period = math.lcm(K - 1, 15)
    T = max(1, round(210 / period)) * period
    acts = torch.randint(0, 28, (T, 4096), device=dev)
    for _ in range(3):
        env.rollout(acts, graph=True)
    env.join(); torch.cuda.synchronize()
    x, y = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = max(2, 2100 // T)
    x.record()
    for _ in range(reps):
        env.rollout(acts, graph=True)
    env.join(); y.record(); torch.cuda.synchronize()
    steady = x.elapsed_time(y) * 1e3 / (reps * T)
    short = []
    for _ in range(15):
        env.rollout(acts[:5]); env.join(); torch.cuda.synchronize()
        torch.cuda._sleep(200_000)
        x.record(); env.rollout(acts[:20]); env.join(); y.record(); torch.cuda.synchronize()
        short.append(x.elapsed_time(y) * 1e3 / 20)
    short.sort()
    print(json.dumps({"ring": K, "chunk": T, "steady_us": round(steady, 2), "short20_us_median": round(short[7], 2), "err": env.error_word()}))
else:
    for K in (8, 16, 32, 64, 8, 32):
        r = subprocess.run([sys.executable, os.path.abspath(__file__), "child", str(K)], capture_output=True, text=True)
        print(r.stdout.strip() or r.stderr[-500:], flush=True)
