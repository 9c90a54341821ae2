"""Host-buffer step (FFMPVectorEnv.step_host) time of the bench workload for (slots, regen_batch, FFMP_TICK_PDL)."""
import json, os, subprocess, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if len(sys.argv) > 1 and sys.argv[1] == "child":
    import torch
    import flow_field_based_motion_planner_b200 as ffmp
    S, m = int(sys.argv[2]), int(sys.argv[3])
    env = ffmp.FFMPVectorEnv(4096, grid=128, window=100, seed=1234, slots=S, regen_batch=m)
    env.reset()
    host_actions = [torch.randint(0, 28, (4096,), dtype=torch.int64).pin_memory() for _ in range(16)]
    for i in range(200):
        env.step_host(host_actions[i % 16])
    env.join(); torch.cuda.synchronize()
    res = []
    for k in (2000, 20, 20, 20, 20, 20):
        t0 = time.perf_counter()
        for i in range(k):
            env.step_host(host_actions[i % 16])
        env.join(); torch.cuda.synchronize()
        res.append(round((time.perf_counter() - t0) / k * 1e6, 2))
    print(json.dumps({"S": S, "m": m, "pdl": os.environ.get("FFMP_TICK_PDL", "1"), "us_per_step_2000": res[0], "us_per_step_20": res[1:]}))
else:
    for S, m, pdl in ((8, 1, "0"), (8, 1, "1"), (16, 1, "1"), (16, 3, "1"), (16, 3, "0"), (16, 2, "1"), (8, 1, "0"), (16, 3, "1")):
        r = subprocess.run([sys.executable, os.path.abspath(__file__), "child", str(S), str(m)],
                           env=dict(os.environ, FFMP_TICK_PDL=pdl), capture_output=True, text=True)
        print(r.stdout.strip() or r.stderr[-400:], flush=True)
