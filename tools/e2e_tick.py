import os, sys, time, torch
sys.path.insert(0, "/root/repo")
import flow_field_based_motion_planner_b200 as ffmp
N = 4096
for p_occ in (0.10, 0.0):
    env = ffmp.FFMPVectorEnv(N, grid=128, window=100, seed=1234, p_occ=p_occ)
    env.reset()
    acts = [torch.randint(0, 28, (N,), dtype=torch.int64).pin_memory() for _ in range(16)]
    dacts = torch.randint(0, 28, (250, N), device="cuda")
    for i in range(300):
        env.step_host(acts[i % 16])
    env.kernel_timing(True)
    t0 = time.perf_counter()
    for i in range(250):
        env.step_host(acts[i % 16])
    dt = time.perf_counter() - t0
    kt = env.kernel_timing(False)
    env.join(); torch.cuda.synchronize()
    env.kernel_timing(True)
    env.rollout(dacts)
    kr = env.kernel_timing(False)
    env.join(); torch.cuda.synchronize()
    print(f"p_occ={p_occ}: e2e {dt/250*1e6:.1f} us/step; tick+export in e2e mode {kt['tick_ms']*1e3:.1f} us (regen launch {kt['regen_ms']*1e3:.0f} us); "
          f"tick in rollout mode {kr['tick_ms']*1e3:.1f} us (regen {kr['regen_ms']*1e3:.0f} us); dones/step {float(env.done.float().sum()):.0f}")
    env.close()
