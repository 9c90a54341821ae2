"""Kernel micro-benchmarks (CUDA events, current stream) used while tuning; prints one JSON object.

    python tools/kbench.py [--envs 4096] [--grid 128] [--window 100]
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import flow_field_based_motion_planner_b200 as ffmp  # noqa: E402


def timed(fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    out = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        out.append(a.elapsed_time(b))
    out.sort()
    return out[len(out) // 2], out[0]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=4096)
    ap.add_argument("--grid", type=int, default=128)
    ap.add_argument("--window", type=int, default=100)
    ap.add_argument("--ring", type=int, default=8)
    ap.add_argument("--slots", type=int, default=6)
    ap.add_argument("--steps", type=int, default=2000)
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    N, G, W = a.envs, a.grid, a.window
    res = {"envs": N, "grid": G, "window": W}
    L = ffmp.native.lib()

    gids = torch.arange(N, device=dev)
    eps = torch.zeros_like(gids)
    occ, scen = ffmp.ops.generate_scenarios(gids, eps, G, seed=1)
    med, best = timed(lambda: ffmp.ops.generate_scenarios(gids, eps, G, seed=1))
    res["scenario_ms"] = med
    goals = scen[:, 5:7].contiguous()
    med, best = timed(lambda: ffmp.ops.flow_field(occ, goals))
    res["flow_field_ms"], res["flow_field_best_ms"] = med, best
    res["flow_field_cells_per_s"] = N * G * G / (med * 1e-3)
    res["flow_field_frac_6B"] = N * G * G * 6 / (med * 1e-3) / 6549.8e9
    med, _ = timed(lambda: ffmp.ops.flow_field(occ, goals, want_cost=False))
    res["flow_field_nocost_ms"] = med
    # single-env latency (what the background regeneration sees when few envs finish)
    med, _ = timed(lambda: ffmp.ops.flow_field(occ[:64], goals[:64]))
    res["flow_field_64env_ms"] = med

    # step kernel without any episode end: no obstacles, huge max_steps, action 3 (stand still)
    env = ffmp.FFMPVectorEnv(N, grid=G, window=W, ring=a.ring, slots=a.slots, p_occ=0.0, max_steps=10 ** 9, seed=3)
    env.reset()
    still = torch.full((100, N), 3, dtype=torch.int64, device=dev)
    med, best = timed(lambda: env.rollout(still), reps=5)
    res["step_only_us"] = med * 10.0          # per step (100 steps per call)
    res["step_only_frac"] = N * (2 * W * W + 146) / (med * 1e-5) / 6549.8e9
    env.close()

    # the bench workload
    env = ffmp.FFMPVectorEnv(N, grid=G, window=W, ring=a.ring, slots=a.slots, seed=1234)
    env.reset()
    for dt, nm, wb in ((torch.float32, "f32", 4), (torch.bfloat16, "bf16", 2)):
        out = torch.empty((N, 2, W, W), dtype=dt, device=dev)
        med, _ = timed(lambda: env.learner_input(dtype=dt, out=out))
        res[f"learner_input_{nm}_us"] = med * 1e3
        res[f"learner_input_{nm}_frac"] = N * 2 * W * W * (1 + wb) / (med * 1e-3) / 6549.8e9
    # LiDAR scan synthesis (SPEC §9) of every env at its current pose: 360 beams, 3.5 m
    sc_out = torch.empty((N, 360), dtype=torch.float32, device=dev)
    med, best = timed(lambda: env.scan(360, 3.5, out=sc_out))
    res["scan360_us"], res["scan360_best_us"] = med * 1e3, best * 1e3
    sc = env.scan(360, 3.5, out=sc_out)[0]
    fin = torch.isfinite(sc)
    res["scan360_beams_per_s"] = N * 360 / (med * 1e-3)
    res["scan360_mean_range_m"] = float(sc[fin].mean().item())
    res["scan360_no_return_frac"] = float((~fin).float().mean().item())
    # device replay ring: one push = the packed transition block of the whole batch
    ring = ffmp.ReplayRing(env, 4)
    med, _ = timed(lambda: ring.push())
    res["replay_push_us"] = med * 1e3
    res["replay_push_frac"] = 2 * ring.block / (med * 1e-3) / 6549.8e9
    del ring
    acts = torch.randint(0, 28, (250, N), device=dev)
    for _ in range(2):
        env.rollout(acts)
    env.join()
    torch.cuda.synchronize()
    x, y = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    x.record()
    k = 0
    while k < a.steps:
        env.rollout(acts)
        k += 250
    env.join()
    y.record()
    torch.cuda.synchronize()
    ms = x.elapsed_time(y)
    res["bench_us_per_step"] = ms * 1e3 / k
    res["bench_env_steps_per_s"] = N * k / (ms * 1e-3)
    res["dones_per_step"] = float(env.done.float().mean().item()) * N
    env.close()
    print(json.dumps(res))


if __name__ == "__main__":
    main()
