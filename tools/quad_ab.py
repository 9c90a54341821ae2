"""Four-warps-per-grid regeneration kernel (flow_field_quad_kernel) against the warp-per-grid kernel (FFMP_FLOW_QUAD=0) on the
bench workload: steady-state step, 20-step window (which pays the join of the last regenerations), the regeneration launch
itself, the host-buffer step, and the reset latency of a single env (the reference's own use)."""
import json, os, subprocess, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if len(sys.argv) > 1 and sys.argv[1] == "child":
    import torch
    import flow_field_based_motion_planner_b200 as ffmp
    dev = torch.device("cuda:0")
    S, m = int(sys.argv[2]), int(sys.argv[3])
    env = ffmp.FFMPVectorEnv(4096, grid=128, window=100, seed=1234, slots=S, regen_batch=m)
    env.reset()
    acts = torch.randint(0, 28, (210, 4096), device=dev)
    for _ in range(2):
        env.rollout(acts)
    env.join(); torch.cuda.synchronize()
    x, y = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    x.record()
    for _ in range(8):
        env.rollout(acts)
    env.join(); y.record(); torch.cuda.synchronize()
    steady = x.elapsed_time(y) * 1e3 / (8 * 210)
    short = []
    for _ in range(20):
        env.rollout(acts[:5]); env.join(); torch.cuda.synchronize()
        x.record(); env.rollout(acts[:20]); env.join(); y.record(); torch.cuda.synchronize()
        short.append(x.elapsed_time(y) * 1e3 / 20)
    short.sort()
    env.kernel_timing(True); env.rollout(acts); kt = env.kernel_timing(False); env.join(); torch.cuda.synchronize()
    # host-buffer steps
    ha = torch.randint(0, 28, (200, 4096)).pin_memory()
    for i in range(50):
        env.step_host(ha[i])
    t0 = time.perf_counter()
    for i in range(200):
        env.step_host(ha[i])
    e2e = (time.perf_counter() - t0) * 1e6 / 200
    err = env.error_word()
    torch.cuda.synchronize(); t0 = time.perf_counter(); env.reset(); torch.cuda.synchronize()
    reset_ms = (time.perf_counter() - t0) * 1e3
    env.close()
    one = ffmp.FFMPVectorEnv(1, grid=128, window=100, seed=5)
    one.reset(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(50):
        one.reset(); torch.cuda.synchronize()
    reset1 = (time.perf_counter() - t0) * 1e6 / 50
    print(json.dumps({"quad": os.environ.get("FFMP_FLOW_QUAD", "default"), "S": S, "m": m, "steady_us": round(steady, 2),
                      "short20_us_median": round(short[10], 2), "short20_us_min": round(short[0], 2),
                      "tick_us": round(kt["tick_ms"] * 1e3, 2), "regen_launch_us": round(kt["regen_ms"] * 1e3, 2),
                      "step_host_us": round(e2e, 2), "reset_1env_us": round(reset1, 1), "reset_4096env_ms": round(reset_ms, 2), "err": err}))
else:
    runs = [({"FFMP_FLOW_QUAD": "0"}, 16, 3), ({}, 16, 3), ({"FFMP_FLOW_QUAD": "4"}, 16, 3), ({"FFMP_FLOW_QUAD": "0"}, 8, 1), ({}, 8, 1)]
    for e, S, m in runs:
        r = subprocess.run([sys.executable, os.path.abspath(__file__), "child", str(S), str(m)],
                           env=dict(os.environ, **e), capture_output=True, text=True)
        print(json.dumps(e), r.stdout.strip() or r.stderr[-600:], flush=True)
