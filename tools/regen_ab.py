"""Step time of the bench workload vs the size of the background regeneration launch (development switch FFMP_REGEN_GRID)."""
import json, os, subprocess, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if len(sys.argv) > 1 and sys.argv[1] == "child":
    import torch
    import flow_field_based_motion_planner_b200 as ffmp
    dev = torch.device("cuda:0")
    env = ffmp.FFMPVectorEnv(4096, grid=128, window=100, seed=1234)
    env.reset()
    acts = torch.randint(0, 28, (250, 4096), device=dev)
    for _ in range(2):
        env.rollout(acts)
    env.join(); torch.cuda.synchronize()
    x, y = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    x.record()
    for _ in range(8):
        env.rollout(acts)
    env.join(); y.record(); torch.cuda.synchronize()
    env.kernel_timing(True); env.rollout(acts); kt = env.kernel_timing(False); env.join(); torch.cuda.synchronize()
    print(json.dumps({"us_per_step": x.elapsed_time(y) / 2.0, "tick_us": kt["tick_ms"] * 1e3, "regen_us": kt["regen_ms"] * 1e3}))
else:
    for g in (74, 148, 222, 296):
        r = subprocess.run([sys.executable, os.path.abspath(__file__), "child"], env=dict(os.environ, FFMP_REGEN_GRID=str(g)), capture_output=True, text=True)
        print(g, r.stdout.strip() or r.stderr[-300:], flush=True)
