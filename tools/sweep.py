"""Sensitivity sweeps used while tuning (one JSON line per case): scenario-slot depth of the bench workload, and
the other BASELINE.json shapes (config 2: 1024 x 64x64 static goal; config 4: 512 x 512x512 dense maps)."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import flow_field_based_motion_planner_b200 as ffmp  # noqa: E402

dev = torch.device("cuda:0")
torch.manual_seed(1234)


def run_env(tag, steps=2000, **kw):
    env = ffmp.FFMPVectorEnv(**kw)
    N = env.num_envs
    env.reset()
    acts = torch.randint(0, 28, (250, N), device=dev)
    for _ in range(2):
        env.rollout(acts)
    env.join()
    torch.cuda.synchronize()
    x, y = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    x.record()
    k = 0
    while k < steps:
        env.rollout(acts)
        k += 250
    env.join()
    y.record()
    torch.cuda.synchronize()
    ms = x.elapsed_time(y)
    env.kernel_timing(True)
    env.rollout(acts)
    kt = env.kernel_timing(False)
    env.join()
    torch.cuda.synchronize()
    out = {"case": tag, "us_per_step": ms * 1e3 / k, "env_steps_per_s": N * k / (ms * 1e-3), "tick_us": kt["tick_ms"] * 1e3,
           "regen_us": kt["regen_ms"] * 1e3, "dones_per_step": float(env.done.float().mean().item()) * N}
    env.close()
    print(json.dumps(out), flush=True)


def run_flow(tag, n, G, p_occ, bs, reps=3):
    gids = torch.arange(n, device=dev)
    occ, scen = ffmp.ops.generate_scenarios(gids, torch.zeros_like(gids), G, p_occ=p_occ, block_shift=bs, seed=7)
    goals = scen[:, 5:7].contiguous()
    ffmp.ops.flow_field(occ, goals)
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        x, y = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        x.record()
        cost, _ = ffmp.ops.flow_field(occ, goals)
        y.record()
        torch.cuda.synchronize()
        best = min(best, x.elapsed_time(y))
    reach = cost[cost != 0x7FFFFFFF]
    print(json.dumps({"case": tag, "n": n, "grid": G, "ms": best, "cells_per_s": n * G * G / (best * 1e-3),
                      "frac_6B": n * G * G * 6 / (best * 1e-3) / 6549.8e9, "max_depth": int(reach.max().item())}), flush=True)


which = sys.argv[1:] or ["slots", "configs", "flow"]
if "slots" in which:
    for S in (3, 4, 6, 8, 12, 16):
        run_env(f"bench workload, slots={S}", num_envs=4096, grid=128, window=100, slots=S, seed=1234)
if "configs" in which:
    run_env("config2: 1024 x 64x64, W=64, static goal", num_envs=1024, grid=64, window=64, goal_mode=1, seed=1234)
    run_env("config4: 512 x 512x512, p=0.3 bs=0", steps=500, num_envs=512, grid=512, window=100, p_occ=0.3, block_shift=0, slots=3, seed=1234)
    run_env("config4b: 512 x 512x512, p=0.1 bs=3", steps=500, num_envs=512, grid=512, window=100, slots=3, seed=1234)
if "flow" in which:
    run_flow("flow 4096x128 p=.1 bs=3", 4096, 128, 0.1, 3)
    run_flow("flow 64x128 (latency)", 64, 128, 0.1, 3)
    run_flow("flow 1x128 (latency)", 1, 128, 0.1, 3)
    run_flow("flow 1024x64", 1024, 64, 0.1, 3)
    run_flow("flow 512x512 p=.3 bs=0", 512, 512, 0.3, 0)
    run_flow("flow 512x512 p=.1 bs=3", 512, 512, 0.1, 3)
    run_flow("flow 148x512 p=.1 bs=3", 148, 512, 0.1, 3)
