"""Host-buffer step (FFMPVectorEnv.step_host) of the bench workload: actions inside the step kernel's launch (by-value
parameter, the default up to 4096 envs) against the copied (FFMP_ACT_PARAM=0) and the in-place (FFMP_HOST_IO=2) forms."""
import json, os, subprocess, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if len(sys.argv) > 1 and sys.argv[1] == "child":
    import torch
    import flow_field_based_motion_planner_b200 as ffmp
    N = int(sys.argv[2])
    env = ffmp.FFMPVectorEnv(N, grid=128, window=100, seed=1234)
    env.reset()
    host_actions = [torch.randint(0, 28, (N,), dtype=torch.int64).pin_memory() for _ in range(16)]
    for i in range(200):
        env.step_host(host_actions[i % 16])
    env.join(); torch.cuda.synchronize()
    res = []
    for k in (2000, 20, 20, 20, 2000):
        t0 = time.perf_counter()
        for i in range(k):
            env.step_host(host_actions[i % 16])
        env.join(); torch.cuda.synchronize()
        res.append(round((time.perf_counter() - t0) / k * 1e6, 2))
    print(json.dumps({"N": N, "us_per_step_2000": [res[0], res[4]], "us_per_step_20": res[1:4], "err": env.error_word()}))
else:
    runs = [({}, 4096), ({"FFMP_ACT_PARAM": "0"}, 4096), ({"FFMP_HOST_IO": "2"}, 4096), ({}, 4096), ({"FFMP_ACT_PARAM": "0"}, 4096),
            ({}, 1024), ({"FFMP_ACT_PARAM": "0"}, 1024), ({}, 256), ({"FFMP_ACT_PARAM": "0"}, 256)]
    for e, N in runs:
        r = subprocess.run([sys.executable, os.path.abspath(__file__), "child", str(N)], env=dict(os.environ, **e), capture_output=True, text=True)
        print(json.dumps(e), r.stdout.strip() or r.stderr[-600:], flush=True)
        if e == {}:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "child", str(N)], env=dict(os.environ, FFMP_HOST_IO_STATS="1", **e), capture_output=True, text=True)
            print("  stats:", r.stderr.strip()[-900:], flush=True)
