"""bench.py — env-steps/s of the gym_ffmp hot path (BASELINE.json metric) on N B200s of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config 2|3|4]

Workload (config.workload), default --config 3: BASELINE.json configs[2] per-GPU shape — 4096 envs/GPU x 128x128 grids,
W=100 local maps, goal re-sampled every reset (flow-field recompute per reset), uniform random actions.
--config 2: 1024 envs x 64x64 grids, W=64, static goal (BASELINE configs[1]); --config 4: 512 envs IN TOTAL x 512x512
grids with dense i.i.d. obstacles (BASELINE configs[3]; strong scaling: 512 / N envs per GPU).
The headline is the MEDIAN of --windows (5) timed windows of exactly --steps steps each (every window in `ms_windows`).
One "step" = one batched env step of every env on every GPU.  Envs shard by global id with no data-path collective
(SPEC.md §3); the only collective is the max-over-ranks of the timing.  The default line also carries short runs of
configs 2 and 4 (`other_configs`) and a 2000-step steady-state run (`steady_state`) as extras.
"""
import argparse
import json
import math
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "env-steps/sec (flow-field+dynamics)"
UNIT = "env-steps/s"
HBM_FALLBACK_GBS = 6650.0     # /opt/skills/guides/B200_PROFILING.md fallback


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20000)
    ap.add_argument("--warmup", type=int, default=500)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", type=int, default=3, choices=[2, 3, 4], help="BASELINE.json configs[] shape (1-based as in SURVEY 8d)")
    ap.add_argument("--envs", type=int, default=None, help="envs per GPU (default: by --config)")
    ap.add_argument("--grid", type=int, default=None)
    ap.add_argument("--window", type=int, default=None)
    ap.add_argument("--ring", type=int, default=32, help="observation frame slots per env: on a ring wrap (every ring-1 steps) every env "
                    "rewrites its older frame too; 8 -> 32 slots takes the steady step from 20.2 to 19.3 us (profiles/r02f_ring_ab.txt)")
    ap.add_argument("--slots", type=int, default=None)
    ap.add_argument("--goal-mode", type=int, default=None)
    ap.add_argument("--p-occ", type=float, default=None)
    ap.add_argument("--block-shift", type=int, default=None)
    ap.add_argument("--seed", type=int, default=1234)
    ap.add_argument("--chunk", type=int, default=None, help="steps per rollout call (the action block is reused); default: the "
                    "multiple of the ring / regeneration-list period nearest to 210, so that every call replays one graph")
    ap.add_argument("--windows", type=int, default=5, help="K-step windows timed back to back (each from reset + warm-up); the median is reported")
    ap.add_argument("--no-graph", action="store_true", help="plain kernel launches (ffmp_rollout) instead of graph replays")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the flow-field / e2e / roofline side measurements")
    a = ap.parse_args()
    return apply_config(a, int(os.environ.get("WORLD_SIZE", str(max(1, a.gpus)))))


# SURVEY.md §8(d) synthetic inputs of the BASELINE.json configurations
CONFIGS = {
    2: dict(envs=1024, grid=64, window=64, slots=16, goal_mode=1, p_occ=0.10, block_shift=3, scaling="weak",
            name="BASELINE configs[1]: 1024 envs x 64x64 grid, W=64, static goal"),
    3: dict(envs=4096, grid=128, window=100, slots=16, goal_mode=0, p_occ=0.10, block_shift=3, scaling="weak",
            name="BASELINE configs[2] per-GPU shape: 4096 envs/GPU x 128x128 grid, W=100, goal re-sampled each reset"),
    4: dict(envs=512, grid=512, window=100, slots=3, goal_mode=0, p_occ=0.30, block_shift=0, scaling="strong",
            name="BASELINE configs[3]: 512 envs in total x 512x512 grid, dense i.i.d. obstacles (p=0.3), W=100"),
}


def apply_config(a, world):
    c = CONFIGS[a.config]
    for k in ("envs", "grid", "window", "slots", "goal_mode", "p_occ", "block_shift"):
        if getattr(a, k) is None:
            setattr(a, k, c[k])
    a.scaling = c["scaling"]
    if a.scaling == "strong" and world > 1:
        a.envs = max(1, a.envs // world)         # config 4: 512 envs in total, sharded over the GPUs
    a.config_name = c["name"]
    return a


def workload_config(a, n_gpus):
    return {"workload": f"{a.config_name}; run as {a.envs} envs/GPU x {a.grid}x{a.grid} grid, W={a.window}, "
                        f"goal {'re-sampled each reset' if a.goal_mode == 0 else 'static'}, uniform random actions",
            "baseline_config": a.config,
            "envs_per_gpu": a.envs, "grid": a.grid, "window": a.window, "ring": a.ring, "slots": a.slots,
            "regen_batch": max(1, min(4, (a.slots - 1) // 5)),       # library default (ffmp_b200.h: ticks per regeneration launch)
            "launch": "plain kernel launches" if a.no_graph else "one CUDA graph replay per rollout call (ffmp_rollout_graphed)",
            "p_occ": a.p_occ, "block_shift": a.block_shift, "max_steps": 200, "global_envs": a.envs * n_gpus,
            "parallelism": f"env-sharded x{n_gpus}, no data-path collective",
            "l2": "no explicit flush: resident inputs (flow planes + frame ring) exceed the 126 MB L2"}


def pin_rank_to_cores(local, world):
    """One disjoint slice of the allowed host cores per rank: with all ranks free to roam the same cores, the launch threads
    of 8 ranks interfere (SCALE_r01: 0.92 weak-scaling efficiency with no collective in the step).  Returns the slice."""
    try:
        allowed = sorted(os.sched_getaffinity(0))
        per = len(allowed) // max(1, world)
        if world > 1 and per >= 1:
            mine = allowed[local * per:(local + 1) * per]
            os.sched_setaffinity(0, mine)
            return mine
        return allowed
    except (AttributeError, OSError):
        return []


# ------------------------------------------------------------------------------------------------------
# clocks sampling (recipe: B200_PROFILING.md)
# ------------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.path = tempfile.mktemp(prefix="ffmp_clocks_", suffix=".csv")
        self.proc = None
        self.gpu_index = gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu_index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except OSError:
            self.proc = None

    def stop(self, settle=0.15):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(settle)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        try:
            for line in open(self.path):
                f = [x.strip() for x in line.split(",")]
                if len(f) < 9:
                    continue
                try:
                    sm.append(float(f[1])); smax.append(float(f[2]))
                except ValueError:
                    continue
                for name, val in zip(names, f[5:9]):
                    if val.lower().startswith("active"):
                        reasons.add(name)
            os.unlink(self.path)
        except OSError:
            pass
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------------
# CPU legs: the oracle port on the host cores (test infrastructure used as the *baseline*, never shipped)
# ------------------------------------------------------------------------------------------------------
def cpu_port_throughput(a, envs_per_thread, steps, warmup, threads, repeats=1):
    """env-steps/s of the C oracle (oracle/ffmp_oracle.c) with `threads` host threads, each stepping its own
    shard of `envs_per_thread` envs of the bench workload (ctypes releases the GIL).  repeats > 1: the `steps`-step window
    is timed that many times back to back and the median is returned (a 20-step window of a 16-thread run is ~20 ms)."""
    import numpy as np
    import oracle
    shards = [oracle.OracleVectorEnv(envs_per_thread, grid=a.grid, window=a.window, goal_mode=a.goal_mode,
                                     p_occ=a.p_occ, seed=a.seed, env_id_base=i * envs_per_thread,
                                     block_shift=a.block_shift) for i in range(threads)]
    rng = np.random.default_rng(a.seed)
    acts = rng.integers(0, 28, (warmup + steps * repeats, threads, envs_per_thread))
    for s in shards:
        s.reset()

    def run(lo, hi):
        def work(i):
            for t in range(lo, hi):
                shards[i].step(acts[t, i])
        th = [threading.Thread(target=work, args=(i,)) for i in range(threads)]
        t0 = time.perf_counter()
        for x in th:
            x.start()
        for x in th:
            x.join()
        return time.perf_counter() - t0

    run(0, warmup)
    dts = sorted(run(warmup + r * steps, warmup + (r + 1) * steps) for r in range(repeats))
    dt = dts[len(dts) // 2]
    for s in shards:
        s.close()
    return threads * envs_per_thread * steps / dt, dt


def run_reference(a):
    """--impl reference: the reference's CPU implementation of the path.  The reference is Python with the path's
    flow field / physics / scenario generation living in un-vendored ROS nodes, so nothing compiles
    into oracle/_ref; the timed code is the oracle port (kind "port") on all host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    envs_per_thread = 32 if a.grid <= 128 else 1
    steps = max(1, min(a.steps, 200 if a.grid <= 128 else 20))
    warmup = max(3, min(a.warmup, 10 if a.grid <= 128 else 3))
    repeats = 7 if a.grid <= 128 else 1
    v, dt = cpu_port_throughput(a, envs_per_thread, steps, warmup, cores, repeats)
    sample = (f"{cores} threads x {envs_per_thread} envs x {steps} steps of the same workload (C oracle port, gcc -O2)"
              + (f", median of {repeats} back-to-back windows" if repeats > 1 else ""))
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": a.gpus, "steps": steps,
            "warmup": warmup, "ms_per_step": dt / steps * 1e3, "higher_is_better": True, "scaling": a.scaling,
            "vs_baseline": None, "dtype": "f32+u8", "data": "synthetic", "config": workload_config(a, a.gpus),
            "sample_envs": cores * envs_per_thread,      # the bounded sample actually stepped (config.envs_per_gpu names the workload)
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------------
def run_steps(env, actions, k, chunk, graph):
    done_steps = 0
    while done_steps < k:
        t = min(chunk, k - done_steps)
        env.rollout(actions[:t], graph=graph)
        done_steps += t


def timed_rollout(torch, env, actions, steps, chunk, barrier, graph):
    """`steps` device-resident env steps (rollout calls of `chunk` steps: one CUDA graph replay each, or plain launches
    with --no-graph), CUDA events on the launching stream, the join of every queued background regeneration inside the
    timed region.  Returns (ms, kernels launched inside the region)."""
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    launches0 = env.launch_count()
    # a fixed ~100 us delay kernel in front of the first event: the host queues the whole region while the GPU is busy, so
    # the events bracket device time only (with an idle GPU the host latency of the first launch would be timed as well)
    torch.cuda._sleep(200_000)
    e0.record()
    run_steps(env, actions, steps, chunk, graph)
    env.join()                    # the timed region ends only when every queued regeneration has finished
    e1.record()
    launches = env.launch_count() - launches0
    barrier()
    return e0.elapsed_time(e1), launches


def side_run(torch, ffmp, dev, cfg, steps, warmup, seed):
    """A short device-resident run of another BASELINE configuration on this GPU (extras of the default line)."""
    env = ffmp.FFMPVectorEnv(cfg["envs"], grid=cfg["grid"], window=cfg["window"], slots=cfg["slots"], goal_mode=cfg["goal_mode"],
                             p_occ=cfg["p_occ"], block_shift=cfg["block_shift"], seed=seed, device=str(dev))
    N = cfg["envs"]
    actions = torch.randint(0, 28, (min(steps, 250), N), device=dev, dtype=torch.int64)
    env.reset()
    env.rollout(actions[:max(3, warmup)])
    env.join()
    torch.cuda.synchronize()
    ms, _ = timed_rollout(torch, env, actions, steps, actions.shape[0], torch.cuda.synchronize, False)
    out = {"workload": cfg["name"], "envs": N, "grid": cfg["grid"], "window": cfg["window"], "steps": steps,
           "us_per_step": ms * 1e3 / steps, "env_steps_per_s": N * steps / (ms * 1e-3),
           "dones_per_step": float(env.done.float().mean().item()) * N}
    env.close()
    del env
    torch.cuda.empty_cache()
    return out


def run_ours(a):
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    cores = pin_rank_to_cores(local, world)      # before CUDA creates its threads

    import torch
    import torch.distributed as dist

    import flow_field_based_motion_planner_b200 as ffmp

    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    N = a.envs

    env = ffmp.FFMPVectorEnv(N, grid=a.grid, window=a.window, ring=a.ring, slots=a.slots, goal_mode=a.goal_mode,
                             p_occ=a.p_occ, block_shift=a.block_shift, seed=a.seed, env_id_base=rank * N, device=f"cuda:{local}")
    gen = torch.Generator(device=dev)
    gen.manual_seed(a.seed + rank)
    graph = not a.no_graph
    regen_batch = max(1, min(4, (a.slots - 1) // 5))
    period = math.lcm(max(1, a.ring - 1), ((a.slots - 1) // regen_batch) * regen_batch)    # ring phase x regeneration-list phase
    long_chunk = a.chunk or max(1, round(210 / period)) * period
    chunk = max(1, min(long_chunk, a.steps))
    actions = torch.randint(0, 28, (max(chunk, long_chunk), N), generator=gen, device=dev, dtype=torch.int64)
    W = max(3, a.warmup)

    def start_and_warm():
        env.reset()
        run_steps(env, actions, W, chunk, graph)
        env.join()

    if graph:
        # graph replays: the first pass through a (steps, ring phase, list phase) combination captures and instantiates
        # its graph (milliseconds).  reset() makes the phases deterministic, so one dry pass of the very same sequence
        # leaves every graph of the timed region in the cache; the timed region then only launches them.
        for _ in range(2):
            start_and_warm()
            run_steps(env, actions, a.steps, chunk, graph)
            env.join()
            torch.cuda.synchronize()
    env.reset()
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world > 1:
            t = torch.tensor([x], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())
        return x

    # ---- headline: K device-resident steps, CUDA events, max over ranks.  The window (reset -> W warm-up steps -> barrier ->
    #      EXACTLY K timed steps + join -> barrier) is run `--windows` times and the MEDIAN window is reported: a single 0.4 ms
    #      window is at the mercy of whatever else touches the GPU in that moment (between boxes and runs it read 0.41-0.48 ms,
    #      inside one process 0.410-0.425 ms, profiles/r02f_window_repeats.txt); every window's time is in `ms_windows`.
    sampler = ClockSampler(local)
    if rank == 0 and not os.environ.get("BENCH_NO_SAMPLER"):
        sampler.start()                       # nvidia-smi needs ~0.1 s to come up: started ahead of the windows, not inside one
    def keep_busy(seconds):
        # untimed rollouts of the same workload: the clock samples bracket the windows under the load they run at
        t_end = time.perf_counter() + seconds
        while time.perf_counter() < t_end:
            env.rollout(actions[:chunk])
            torch.cuda.synchronize()

    keep_busy(0.3)
    start_and_warm()
    windows = []
    for w in range(max(1, a.windows)):
        if w > 0:
            start_and_warm()
        barrier()
        m, timed_launches = timed_rollout(torch, env, actions, a.steps, chunk, barrier, graph)
        if world > 1:
            t = torch.zeros(world, device=dev, dtype=torch.float64)
            t[rank] = m
            dist.all_reduce(t)
            per_rank = [round(float(x), 4) for x in t.tolist()]
        else:
            per_rank = [round(m, 4)]
        windows.append((max(per_rank), per_rank))
    keep_busy(0.15)
    clocks = sampler.stop(settle=0.0) if rank == 0 else None
    order = sorted(range(len(windows)), key=lambda i: windows[i][0])
    ms, ms_per_rank = windows[order[len(order) // 2]]
    value = world * N * a.steps / (ms * 1e-3)

    extras = {"ms_windows": [w[0] for w in windows]}
    e2e = None
    roofline = None
    if not a.no_extras:
        # ---- e2e: public Python API with HOST buffers (pinned actions in, reward/done/goal/velocity out) ----
        k2 = max(10, min(a.steps, 2000))
        host_actions = [torch.randint(0, 28, (N,), dtype=torch.int64).pin_memory() for _ in range(16)]
        # warm-up: in a fresh process the host-buffer step takes ~37 us for its first ~500 calls (~20 ms) and 32.5 us from then
        # on, whatever ran on the GPU before (profiles/r02f_e2e_settling.txt), so 1000 untimed steps precede the timed ones
        for i in range(1000):
            env.step_host(host_actions[i % 16])
        # like the headline: `--windows` windows of exactly k2 calls (+ join + synchronize), the median window is reported
        e2e_windows = []
        for w in range(max(1, a.windows if k2 < 2000 else 1)):
            for i in range(50 if w else 0):
                env.step_host(host_actions[i % 16])
            barrier()
            t0 = time.perf_counter()
            for i in range(k2):
                env.step_host(host_actions[i % 16])
            t1 = time.perf_counter()
            env.join()
            torch.cuda.synchronize()
            t2 = time.perf_counter()
            e2e_windows.append((max_over_ranks(t2 - t0), t0, t1, t2))
        dt, t0, t1, t2 = sorted(e2e_windows)[len(e2e_windows) // 2]
        e2e = {"value": world * N * k2 / dt, "unit": UNIT, "h2d_bytes_per_step": env.h2d_bytes_per_step * world,
               "d2h_bytes_per_step": env.d2h_bytes_per_step * world, "steps": k2,
               "windows_env_steps_per_s": [round(world * N * k2 / x[0]) for x in e2e_windows],
               "us_per_step_calls_only": (t1 - t0) * 1e6 / k2, "join_and_sync_us": (t2 - t1) * 1e6,
               "note": "FFMPVectorEnv.step_host every step, host buffers in and out: the caller's pinned int64 actions are "
                       "narrowed to one byte per env on the host and ride inside the step kernel's launch as a by-value "
                       "parameter (up to 4096 envs per GPU; a cudaMemcpyAsync above that), reward/done/flags/relative_goal/velocity "
                       "are written into the caller's pinned block by a kernel queued behind the step, and the call returns when "
                       "its completion word (mapped memory) is set: no copy engine in either direction, no stream sync; "
                       "h2d_bytes_per_step counts the caller's int64 buffer; the 20 KB/env local_map observations stay on the "
                       "device for the learner"}

        # ---- steady state: the same step over a long window (the join of the last regeneration, ~50-90 us, is 15 % of a
        #      20-step window and < 0.5 % of this one) ----
        if a.steps < 2000:
            ss_steps = 10 * long_chunk
            run_steps(env, actions, 2 * long_chunk, long_chunk, graph)      # the chunk's graph at the phase this run starts from
            env.join()
            ss_ms, _ = timed_rollout(torch, env, actions, ss_steps, long_chunk, barrier, graph)
            ss_ms = max_over_ranks(ss_ms)
            extras["steady_state"] = {"steps": ss_steps, "ms_per_step": ss_ms / ss_steps, "value": world * N * ss_steps / (ss_ms * 1e-3),
                                      "unit": UNIT, "chunk": long_chunk,
                                      "note": f"same workload, {ss_steps} timed steps instead of --steps"}

        # ---- roofline of the dominant kernel of the step (tick_tma_kernel: the whole env step in one launch), timed
        #      live with CUDA events recorded by the library on the launching stream around each launch (ffmp_timing);
        #      the same call times the background regeneration (flow-field) launch of each tick on its side stream ----
        if rank == 0:
            peaks = {}
            try:
                peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
            except (OSError, ValueError):
                pass
            peak = float(peaks.get("hbm_gbs", HBM_FALLBACK_GBS))
            prof = {}
            try:
                prof = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
            except (OSError, ValueError):
                pass
            env.join()
            torch.cuda.synchronize()
            env.kernel_timing(True)
            env.rollout(actions[:min(long_chunk, 250)])
            kt = env.kernel_timing(False)
            env.join()
            torch.cuda.synchronize()
            # 2*W^2 + 146 algorithmic bytes per env-step (window read + frame write + state/outputs)
            tick_bytes = N * (2 * a.window * a.window + 146)
            # The launching stream carries nothing but tick launches, each a programmatic dependent of the one before, so
            # the kernel's average launch duration over a timed region is region / launches (events at both ends of the
            # region).  Events around EVERY launch (ffmp_timing) serialise the launches and add their own cost; that
            # figure is kept as `isolated_launch_ms`.
            ss = extras.get("steady_state")
            tick_ms = ss["ms_per_step"] if ss else ms / a.steps
            achieved = tick_bytes / (tick_ms * 1e-3) / 1e9
            roofline = {"kernel": "tick_tma_kernel", "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                        "frac": achieved / peak,
                        "traffic": prof.get("tick_kernel_dram_bytes_per_launch") if a.config == 3 and N == 4096 else None,
                        "traffic_source": prof.get("tick_source", "profiles/traffic.json absent") + " (one launch under ncu: the "
                                          "frame writes of that launch are still dirty in the 126 MB L2, so DRAM traffic reads below "
                                          "the algorithmic bytes)",
                        "peak_source": "MEASURED_PEAKS.json hbm_gbs (measured)" if "hbm_gbs" in peaks else "fallback 6650 GB/s",
                        "algorithmic_bytes_per_launch": tick_bytes, "avg_launch_ms": tick_ms,
                        "launches_timed": ss["steps"] if ss else a.steps,
                        "timing": "timed region / launches: back-to-back launches on the launching stream under the bench "
                                  "workload's background regeneration load" + (" (steady_state run)" if ss else ""),
                        "isolated_launch_ms": kt["tick_ms"], "isolated_launches_timed": kt["ticks"],
                        "isolated_frac": tick_bytes / (kt["tick_ms"] * 1e-3) / 1e9 / peak,
                        "regen_launch_avg_ms": kt["regen_ms"],
                        "step": {"algorithmic_bytes": tick_bytes, "ms": ms / a.steps,
                                 "achieved": tick_bytes / (ms / a.steps * 1e-3) / 1e9,
                                 "frac": tick_bytes / (ms / a.steps * 1e-3) / 1e9 / peak,
                                 "note": "the --steps timed region (incl. the join of the last background regenerations) vs the same bytes"}}

            # ---- flow-field operator: grid cells/s, fraction of the 6 B/cell HBM roofline, and the ALU-pipe ceiling.
            #      10 back-to-back launches per timing with preallocated outputs: the ~30 us Python call of one launch
            #      overlaps the previous launch instead of being timed with an idle GPU in front ----
            FN = N if a.grid <= 128 else min(N, 512)
            gids = torch.arange(FN, device=dev)
            occ, scen = ffmp.ops.generate_scenarios(gids, torch.zeros_like(gids), a.grid, p_occ=a.p_occ, block_shift=a.block_shift,
                                                    seed=a.seed)
            goals = scen[:, 5:7].to(torch.int32).contiguous()
            ws = ffmp.ops.flow_field_workspace(FN, a.grid, dev)
            bufs = (torch.empty((FN, a.grid, a.grid), dtype=torch.int32, device=dev),
                    torch.empty((FN, a.grid, a.grid), dtype=torch.uint8, device=dev))
            reps = 10 if a.grid <= 128 else 2
            for _ in range(3):
                ffmp.ops.flow_field(occ, goals, out=bufs, workspace=ws)
            torch.cuda.synchronize()
            ff = []
            for _ in range(5):
                x, y = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                x.record()
                for _ in range(reps):
                    ffmp.ops.flow_field(occ, goals, out=bufs, workspace=ws)
                y.record()
                torch.cuda.synchronize()
                ff.append(x.elapsed_time(y) / reps)
            ff.sort()
            ff_ms = ff[len(ff) // 2]
            cells = FN * a.grid * a.grid
            extras["flow_field"] = {"cells_per_s": cells / (ff_ms * 1e-3), "ms": ff_ms, "batch": FN, "grid": a.grid,
                                    "algorithmic_bytes_per_cell": 6, "launches_per_timing": reps,
                                    "achieved_gbs": cells * 6 / (ff_ms * 1e-3) / 1e9,
                                    "frac_of_hbm_peak": cells * 6 / (ff_ms * 1e-3) / 1e9 / peak}
            if a.grid == 128 and "flow_il_alu_inst_per_grid" in prof:
                # second roofline of the same kernel: it moves only its algorithmic bytes (ncu: DRAM traffic = 0.87 x
                # algorithmic) and spends its time in logic ops, which issue on the ALU pipe at one warp instruction every
                # other cycle per scheduler.  Counted ALU-pipe / all warp instructions per grid come from the committed
                # ncu capture (profiles/traffic.json), the pipe rate from tools/pipe_probe.cu (profiles/r01a_pipe_probe.txt).
                clk = float((clocks or {}).get("sm_mhz") or peaks.get("sm_max_mhz", 1965.0)) * 1e6
                alu_peak = float(prof.get("alu_pipe_warp_inst_per_clk_per_sm", 1.96)) * 148 * clk
                alu_ach = float(prof["flow_il_alu_inst_per_grid"]) * FN / (ff_ms * 1e-3)
                issue_ach = float(prof["flow_il_inst_per_grid"]) * FN / (ff_ms * 1e-3)
                extras["flow_field"]["roofline_alu"] = {
                    "bound": "alu", "kernel": "flow_field_il_kernel", "achieved": alu_ach / 1e9, "peak": alu_peak / 1e9,
                    "unit": "G warp-inst/s (ALU pipe)", "frac": alu_ach / alu_peak,
                    "issue_frac": issue_ach / (4.0 * 148 * clk),
                    "alu_inst_per_grid": prof["flow_il_alu_inst_per_grid"], "inst_per_grid": prof["flow_il_inst_per_grid"],
                    "source": prof.get("flow_source", "profiles/traffic.json"),
                    "ms_at_alu_peak": float(prof["flow_il_alu_inst_per_grid"]) * FN / alu_peak * 1e3}

            # ---- LiDAR scan synthesis: 360 beams x 3.5 m per env at the current poses ----
            sc_out = torch.empty((N, 360), dtype=torch.float32, device=dev)
            for _ in range(3):
                env.scan(360, 3.5, out=sc_out)
            torch.cuda.synchronize()
            x, y = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            x.record()
            for _ in range(5):
                env.scan(360, 3.5, out=sc_out)
            y.record()
            torch.cuda.synchronize()
            sc_ms = x.elapsed_time(y) / 5
            extras["scan"] = {"beams": 360, "range_max_m": 3.5, "ms": sc_ms, "beams_per_s": N * 360 / (sc_ms * 1e-3)}
            del occ, scen, bufs, ws, sc_out

            # ---- Q network forward (train.py:231-303) on the tcgen05 kernels: 256 observations straight
            #      from learner_input's bf16 NCHW, random-init weights of the reference's shapes ----
            if a.grid <= 128 and a.window == 100:
                try:
                    B = min(256, N)
                    qn = ffmp.QNetwork(max_batch=B, device=str(dev))
                    g = torch.Generator(device="cpu"); g.manual_seed(a.seed)
                    sd = {}
                    for name, shp in ffmp.qnet.SHAPES.items():
                        fan_in = 1
                        for d in shp[1:]:
                            fan_in *= d
                        sd[name + ".weight"] = (torch.rand(shp, generator=g) * 2 - 1) / math.sqrt(fan_in)
                        sd[name + ".bias"] = (torch.rand(shp[0], generator=g) * 2 - 1) / math.sqrt(fan_in)
                    qn.load_state_dict(sd)
                    obs_m = env.learner_input(dtype=torch.bfloat16)[:B].contiguous()
                    og, ov = env.rel_goal[:B].clone(), env.velocity[:B].clone()
                    ot = torch.full((B, 1), 0.1, device=dev)
                    qout = torch.empty((B, 28), device=dev)
                    for _ in range(2):
                        qn(obs_m, og, ov, ot, out=qout)
                    torch.cuda.synchronize()
                    x, y = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    x.record()
                    for _ in range(3):
                        qn(obs_m, og, ov, ot, out=qout)
                    y.record()
                    torch.cuda.synchronize()
                    q_ms = x.elapsed_time(y) / 3
                    macs = (69 * 69 * 32 * 2048 + 38 * 38 * 64 * 32768 + 31 * 31 * 64 * 4096 + (24 * 24 + 17 * 17 + 100) * 64 * 4096
                            + 6400 * 512 + 512 * 512 + 512 * 29)
                    tfl = 2 * macs * B / (q_ms * 1e-3) / 1e12
                    extras["qnet"] = {"kernel": "qnet_conv1_kernel + qnet_gemm_kernel (tcgen05 / TMEM / TMA)", "batch": B, "ms": q_ms,
                                      "bound": "tensor", "achieved": tfl, "unit": "TFLOP/s", "dtype": "bf16 operands, f32 accumulate",
                                      "peak": peaks.get("bf16_tflops_sustained"),
                                      "frac": tfl / peaks["bf16_tflops_sustained"] if peaks.get("bf16_tflops_sustained") else None,
                                      "gmac_per_sample": macs / 1e9, "finite": bool(torch.isfinite(qout).all().item())}
                    qn.close()
                    del qn, obs_m
                except Exception as ex:      # a side measurement must not lose the headline line
                    extras["qnet"] = {"error": repr(ex)[:200]}

    env.close()
    del env
    torch.cuda.empty_cache()

    # ---- the other BASELINE configurations on this GPU (short runs; the headline stays --config) ----
    if rank == 0 and world == 1 and not a.no_extras and a.config == 3:
        other = {}
        for cid, steps, warm in ((2, 500, 20), (4, 60, 5)):
            try:
                other[f"config{cid}"] = side_run(torch, ffmp, dev, CONFIGS[cid], steps, warm, a.seed)
            except Exception as ex:      # a side measurement must not lose the headline line
                other[f"config{cid}"] = {"error": repr(ex)[:200]}
        extras["other_configs"] = other

    cpu = None
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        # a bounded sample of the same workload, ~10 s of single-thread CPU work
        if a.grid <= 128:
            v, dt = cpu_port_throughput(a, 256, 2400 if a.grid == 128 else 6000, 20, 1)
            sample = f"256 envs x {2400 if a.grid == 128 else 6000} steps"
        else:
            v, dt = cpu_port_throughput(a, 4, 40, 2, 1)
            sample = "4 envs x 40 steps"
        cpu = {"value": v, "unit": UNIT, "cores": 1, "kind": "port",
               "sample": f"{sample} of the same workload, single thread, {dt:.1f} s (C oracle port, gcc -O2)"}
        # the same port's flow field alone: queue BFS + 8-neighbour argmin on generated maps
        import oracle
        maps = [oracle.scenario(a.seed, k, 0, a.grid, p_occ=a.p_occ, block_shift=a.block_shift) for k in range(64 if a.grid <= 128 else 4)]
        t0 = time.perf_counter()
        reps = 0
        while time.perf_counter() - t0 < 2.0:
            for occ_k, _, _, cells_k in maps:
                oracle.flow_field(occ_k, cells_k[2], cells_k[3])
            reps += 1
        cpu["flow_field_cells_per_s"] = reps * len(maps) * a.grid * a.grid / (time.perf_counter() - t0)
        # the reference's own Python path, timed in the build container where /root/reference exists
        try:
            cpu["python_reference"] = json.load(open(os.path.join(ROOT, "tests", "golden", "python_reference_timing.json")))
        except (OSError, ValueError):
            pass

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": max(3, a.warmup),
                "ms_per_step": ms / a.steps, "ms_per_rank": ms_per_rank, "higher_is_better": True, "scaling": a.scaling, "vs_baseline": None,
                "dtype": "f32+u8", "data": "synthetic", "config": workload_config(a, world), "clocks": clocks,
                "e2e": e2e, "gpu_launches": timed_launches * world, "roofline": roofline, "cpu_baseline": cpu,
                "host_cores_per_rank": len(cores)}
        line.update(extras)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)


if __name__ == "__main__":
    main()
