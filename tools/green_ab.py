"""Experiment: background regeneration confined to an SM partition (green context, FFMP_REGEN_SMS) vs free to roam.
Reports the device-resident rollout step and the host-buffer step of the bench workload.
The switch lived in ffmp_create for this experiment only (commit 0e8fcc8; result: profiles/r02d_green_ctx_ab.txt, no gain) and
is not in the library any more — check that commit out to re-run."""
import json, os, subprocess, sys, time
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
    host_actions = [torch.randint(0, 28, (4096,), dtype=torch.int64).pin_memory() for _ in range(16)]
    for i in range(200):
        env.step_host(host_actions[i % 16])
    env.join(); torch.cuda.synchronize()
    res = []
    for k in (2000, 20, 20, 20):
        t0 = time.perf_counter()
        for i in range(k):
            env.step_host(host_actions[i % 16])
        env.join(); torch.cuda.synchronize()
        res.append(round((time.perf_counter() - t0) / k * 1e6, 2))
    print(json.dumps({"S": S, "m": m, "regen_sms": os.environ.get("FFMP_REGEN_SMS", "0"), "rollout_steady_us": round(steady, 2),
                      "rollout_20_us_median": round(short[10], 2), "host_step_us_2000": res[0], "host_step_us_20": res[1:],
                      "err": env.error_word()}))
else:
    for S, m, k in ((16, 3, 0), (16, 3, 24), (16, 3, 32), (16, 3, 40), (16, 3, 48), (16, 3, 64), (16, 1, 32), (16, 1, 48), (16, 3, 0)):
        r = subprocess.run([sys.executable, os.path.abspath(__file__), "child", str(S), str(m)],
                           env=dict(os.environ, FFMP_REGEN_SMS=str(k)), capture_output=True, text=True)
        print(r.stdout.strip() or r.stderr[-600:], r.stderr.strip()[-120:] if "green" in r.stderr else "", flush=True)
