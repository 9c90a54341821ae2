"""A/B of the host-buffer step paths (FFMP_HOST_IO = 0 copy engines, 1 mapped results, 2 mapped results + in-place actions).
    python tools/e2e_ab.py [envs] [steps]
One process per mode (the variable is read at ffmp_create)."""
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def child(N, steps):
    import torch
    import flow_field_based_motion_planner_b200 as ffmp
    env = ffmp.FFMPVectorEnv(N, grid=128, window=100, seed=1234)
    env.reset()
    acts = [torch.randint(0, 28, (N,), dtype=torch.int64).pin_memory() for _ in range(16)]
    for i in range(200):
        env.step_host(acts[i % 16])
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(steps):
        env.step_host(acts[i % 16])
    env.join()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    # the same loop without the python wrapper's bookkeeping: raw C-ABI calls
    import ctypes as C
    ptrs = [C.c_void_p(a.data_ptr()) for a in acts]
    st = env._stream()
    t0 = time.perf_counter()
    for i in range(steps):
        env._L.ffmp_step_host(env._h, ptrs[i % 16], *env._host_ptrs, st)
    env.join()
    torch.cuda.synchronize()
    dt2 = time.perf_counter() - t0
    print(json.dumps({"host_io": os.environ.get("FFMP_HOST_IO"), "envs": N, "us_per_step": dt / steps * 1e6,
                      "env_steps_per_s": N * steps / dt, "us_per_step_raw_abi": dt2 / steps * 1e6}))


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "child":
        child(int(sys.argv[2]), int(sys.argv[3]))
    else:
        N = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
        steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3000
        for mode in ("0", "1", "2"):
            env = dict(os.environ, FFMP_HOST_IO=mode)
            subprocess.run([sys.executable, __file__, "child", str(N), str(steps)], env=env, check=False)
