"""Grouped regeneration + tick -> tick programmatic dependent launch (ffmp_cfg.regen_batch): steady-state step time of the
bench workload for (scenario slots S, ticks per regeneration launch m), with the attribute on and off (FFMP_TICK_PDL)."""
import json, os, subprocess, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if len(sys.argv) > 1 and sys.argv[1] == "child":
    import torch
    import flow_field_based_motion_planner_b200 as ffmp
    dev = torch.device("cuda:0")
    S, m = int(sys.argv[2]), int(sys.argv[3])
    env = ffmp.FFMPVectorEnv(4096, grid=128, window=100, seed=1234, slots=S, regen_batch=m)
    env.reset()
    acts = torch.randint(0, 28, (252, 4096), device=dev)
    for _ in range(2):
        env.rollout(acts)
    env.join(); torch.cuda.synchronize()
    x, y = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    x.record()
    for _ in range(8):
        env.rollout(acts)
    env.join(); y.record(); torch.cuda.synchronize()
    steady = x.elapsed_time(y) * 1e3 / (8 * 252)
    short = []
    for _ in range(20):
        env.rollout(acts[:5]); env.join(); torch.cuda.synchronize()
        x.record(); env.rollout(acts[:20]); env.join(); y.record(); torch.cuda.synchronize()
        short.append(x.elapsed_time(y) * 1e3 / 20)
    short.sort()
    print(json.dumps({"S": S, "m": m, "pdl": os.environ.get("FFMP_TICK_PDL", "1"), "steady_us": round(steady, 2),
                      "short20_us_median": round(short[10], 2), "short20_us_min": round(short[0], 2), "err": env.error_word()}))
else:
    for S, m, pdl in ((8, 1, "0"), (8, 1, "1"), (8, 2, "1"), (12, 1, "1"), (12, 2, "1"), (12, 2, "0"), (12, 3, "1"), (16, 2, "1"), (16, 3, "1"),
                      (16, 4, "1"), (16, 5, "1")):
        r = subprocess.run([sys.executable, os.path.abspath(__file__), "child", str(S), str(m)],
                           env=dict(os.environ, FFMP_TICK_PDL=pdl), capture_output=True, text=True)
        print(r.stdout.strip() or r.stderr[-400:], flush=True)
