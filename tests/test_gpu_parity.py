"""GPU parity tests proper: the CUDA path (through the C-ABI) against the CPU oracle on identical
seeded inputs.  Integer / byte work is bit-exact; the fp32 SPEC (fixed operation order, no FMA) makes
poses, rewards and observations bit-exact too, which is stricter than the 1e-5 tolerance of the
north star (asserted as well so a future relaxation stays visible)."""
import os

import numpy as np
import pytest
import torch

import oracle
from maps import special_cases

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

pytestmark = pytest.mark.gpu

INF = oracle.INF


@pytest.fixture(scope="module")
def ffmp(cuda_device):
    import flow_field_based_motion_planner_b200 as pkg
    return pkg


def t2n(t):
    return t.detach().cpu().numpy()


# ---------------------------------------------------------------------------------------------------
# M: scenario generator
# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("G,mode,bs,p", [(128, 0, 3, 0.1), (64, 1, 3, 0.1), (100, 0, 0, 0.1), (32, 0, 2, 0.3),
                                         (128, 1, 4, 0.25), (16, 0, 1, 0.2), (96, 0, 3, 0.1)])
def test_scenarios_bit_exact(ffmp, cuda_device, G, mode, bs, p):
    n = 24
    gids = torch.arange(1000, 1000 + n, device=cuda_device)
    eps = torch.arange(n, device=cuda_device) % 5
    occ, scen = ffmp.ops.generate_scenarios(gids, eps, G, p_occ=p, goal_mode=mode, block_shift=bs, seed=0xABCDEF0123)
    occ, scen = t2n(occ), t2n(scen)
    for k in range(n):
        o, start, goal, cells = oracle.scenario(0xABCDEF0123, 1000 + k, k % 5, G, p_occ=p, goal_mode=mode, block_shift=bs)
        assert np.array_equal(occ[k], o), (k, "occ")
        rec = scen[k]
        assert np.array_equal(rec[0:3].view(np.float32), start)
        assert np.array_equal(rec[3:5].view(np.float32), goal)
        assert rec[5:7].tolist() == cells[2:4].tolist()
        assert np.uint32(rec[7]) == np.uint32(oracle.lib().orc_key(0xABCDEF0123, 1000 + k, k % 5))


# ---------------------------------------------------------------------------------------------------
# F1 + F2: integration field and flow direction
# ---------------------------------------------------------------------------------------------------
def run_flow(ffmp, dev, occs, goals, want_cost=True):
    occ_t = torch.as_tensor(np.stack(occs), device=dev)
    goal_t = torch.as_tensor(np.array(goals, np.int32), device=dev)
    cost, flow = ffmp.ops.flow_field(occ_t, goal_t, want_cost=want_cost)
    torch.cuda.synchronize()
    return (t2n(cost) if want_cost else None), t2n(flow)


@pytest.mark.parametrize("G", [16, 32, 64, 96, 100, 112, 124, 128])
def test_flow_field_special_maps_bit_exact(ffmp, cuda_device, G):
    cases = special_cases(G)
    cost, flow = run_flow(ffmp, cuda_device, [c[1] for c in cases], [c[2] for c in cases])
    for k, (name, occ, goal) in enumerate(cases):
        ec, ed, ef = oracle.flow_field(occ, goal[0], goal[1])
        assert np.array_equal(cost[k], ec), (G, name, "cost", int((cost[k] != ec).sum()))
        assert np.array_equal(flow[k], ef), (G, name, "flow", int((flow[k] != ef).sum()))
        d = np.where(flow[k] == 255, 8, flow[k] // 28)
        assert np.array_equal(d, ed), (G, name, "dir")


@pytest.mark.parametrize("G,bs,p", [(128, 3, 0.1), (128, 0, 0.3), (64, 3, 0.1), (128, 2, 0.35), (100, 3, 0.2)])
def test_flow_field_generated_maps_bit_exact(ffmp, cuda_device, G, bs, p):
    n = 96
    occs, goals = [], []
    for k in range(n):
        o, _, _, cells = oracle.scenario(99, k, 0, G, p_occ=p, block_shift=bs)
        occs.append(o)
        goals.append((cells[2], cells[3]))
    cost, flow = run_flow(ffmp, cuda_device, occs, goals)
    for k in range(n):
        ec, ed, ef = oracle.flow_field(occs[k], *goals[k])
        assert np.array_equal(cost[k], ec), (k, "cost")
        assert np.array_equal(flow[k], ef), (k, "flow")


@pytest.mark.parametrize("G", [160, 256, 448, 512])
def test_flow_field_large_maps_bit_exact(ffmp, cuda_device, G):
    """Large-map path (one CTA per grid): special maps incl. the deep serpentine, and BASELINE config 4's
    shape (512x512, dense i.i.d. obstacles p=0.30)."""
    cases = special_cases(G)
    if G >= 448:
        cases = [c for c in cases if c[0] in ("open", "serpentine", "rooms0", "noise0.35", "sealed_pocket", "goal_blocked")]
    for k in range(4):
        o, _, _, cells = oracle.scenario(77, k, 0, G, p_occ=0.30, block_shift=0)
        cases.append((f"config4_{k}", o, (cells[2], cells[3])))
    cost, flow = run_flow(ffmp, cuda_device, [c[1] for c in cases], [c[2] for c in cases])
    for k, (name, occ, goal) in enumerate(cases):
        ec, ed, ef = oracle.flow_field(occ, goal[0], goal[1])
        assert np.array_equal(cost[k], ec), (G, name, "cost", int((cost[k] != ec).sum()))
        assert np.array_equal(flow[k], ef), (G, name, "flow", int((flow[k] != ef).sum()))
    _, flow2 = run_flow(ffmp, cuda_device, [cases[0][1]], [cases[0][2]], want_cost=False)
    assert np.array_equal(flow2[0], oracle.flow_field(cases[0][1], *cases[0][2])[2])


def test_flow_field_cluster_variant_bit_exact(ffmp, cuda_device, monkeypatch):
    """FFMP_FLOW_CLUSTER=1: maps of 384 < G <= 512 on a thread-block cluster of two CTAs per grid (rows split at 256, the seam row
    exchanged through DSMEM, barrier.cluster per level, cluster-wide convergence vote and hand-out) — BASELINE config 4's
    "multi-CTA wavefront per env".  Same bytes as the one-CTA kernel: special maps (incl. goals on either side of the seam and
    corridors crossing it), generated maps, and a config-4 rollout with its background regeneration."""
    monkeypatch.setenv("FFMP_FLOW_CLUSTER", "1")
    for G in (512, 448):
        cases = special_cases(G)
        cost, flow = run_flow(ffmp, cuda_device, [c[1] for c in cases], [c[2] for c in cases])
        for k, (name, occ, goal) in enumerate(cases):
            ec, ed, ef = oracle.flow_field(occ, goal[0], goal[1])
            assert np.array_equal(cost[k], ec), (G, name, "cost", int((cost[k] != ec).sum()))
            assert np.array_equal(flow[k], ef), (G, name, "flow", int((flow[k] != ef).sum()))
    occs, goals = [], []
    for k in range(7):
        o, _, _, cells = oracle.scenario(9, k, 0, 512, p_occ=0.3, block_shift=0)
        occs.append(np.array(o, copy=True))
        goals.append((cells[2], cells[3]) if k % 2 else (250 + 3 * k, 40 + 60 * k))      # goals next to the seam (rows 255 / 256) too
    for k in range(7):
        occs[k][goals[k][0], goals[k][1]] = 0
    cost, flow = run_flow(ffmp, cuda_device, occs, goals)
    for k in range(7):
        ec, _, ef = oracle.flow_field(occs[k], goals[k][0], goals[k][1])
        assert np.array_equal(cost[k], ec) and np.array_equal(flow[k], ef), k
    rollout_parity(ffmp, 6, 60, seed=4, grid=512, window=100, p_occ=0.30, block_shift=0, slots=3, check_every=30)


def test_rollout_large_map_config4_shape(ffmp):
    """config 4 per-env shape: 512x512 grid with dense i.i.d. obstacles (p=0.30), W=100."""
    rollout_parity(ffmp, 6, 120, seed=4, grid=512, window=100, p_occ=0.30, block_shift=0, slots=3, check_every=60)
    rollout_parity(ffmp, 8, 150, seed=5, grid=256, window=100, check_every=75)


def test_two_devices_in_one_process_large_shared_memory(ffmp):
    """the >48 KB dynamic shared-memory opt-in belongs to the device; one process driving two GPUs through the
    C-ABI must get it on both (G = 256 rows kernel: 66 KB per CTA; W = 212 step kernel: 51 KB per CTA).  Skipped on one GPU."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs in one process")
    for dev in ("cuda:0", "cuda:1", "cuda:0"):
        rollout_parity(ffmp, 4, 24, seed=51, grid=256, window=212, device=dev, max_steps=10, check_every=12)
        rollout_parity(ffmp, 3, 12, seed=52, grid=512, window=100, p_occ=0.3, block_shift=0, slots=3, device=dev, check_every=6)


def test_flow_field_without_cost_output(ffmp, cuda_device):
    occs, goals = [], []
    for k in range(8):
        o, _, _, cells = oracle.scenario(5, k, 0, 128)
        occs.append(o); goals.append((cells[2], cells[3]))
    _, flow = run_flow(ffmp, cuda_device, occs, goals, want_cost=False)
    for k in range(8):
        assert np.array_equal(flow[k], oracle.flow_field(occs[k], *goals[k])[2])


def test_flow_field_empty_batch(ffmp, cuda_device):
    cost, flow = ffmp.ops.flow_field(torch.zeros((0, 64, 64), dtype=torch.uint8, device=cuda_device),
                                     torch.zeros((0, 2), dtype=torch.int32, device=cuda_device))
    assert cost.shape == (0, 64, 64) and flow.shape == (0, 64, 64)


def test_flow_field_full_size_properties(ffmp, cuda_device):
    """BASELINE config 3 size (4096 x 128 x 128): size-independent properties on the device, plus an
    oracle check of a sample."""
    n, G = 4096, 128
    dev = cuda_device
    gids = torch.arange(n, device=dev)
    occ, scen = ffmp.ops.generate_scenarios(gids, torch.zeros_like(gids), G, seed=2024)
    goals = scen[:, 5:7].contiguous()
    cost, flow = ffmp.ops.flow_field(occ, goals)
    torch.cuda.synchronize()
    c = cost.to(torch.int64)
    reach = c != INF
    idx = torch.arange(n, device=dev)
    assert (c[idx, goals[:, 0].long(), goals[:, 1].long()] == 0).all()
    assert (c[occ != 0] == INF).all()
    big = torch.full((n, G + 2, G + 2), INF, dtype=torch.int64, device=dev)
    big[:, 1:-1, 1:-1] = c
    nmin = torch.minimum(torch.minimum(big[:, :-2, 1:-1], big[:, 2:, 1:-1]), torch.minimum(big[:, 1:-1, :-2], big[:, 1:-1, 2:]))
    inner = reach.clone()
    inner[idx, goals[:, 0].long(), goals[:, 1].long()] = False
    assert (nmin[inner] == c[inner] - 1).all()
    lost = (occ == 0) & ~reach
    assert (nmin[lost] == INF).all()
    # flow image: occupied <-> 255, reached non-goal cells always have a direction, others "none"
    assert ((flow == 255) == (occ != 0)).all()
    d = torch.where(flow == 255, torch.full_like(flow, 8), flow // 28)
    assert (d[inner] < 8).all() and (d[~inner] == 8).all()
    # the chosen neighbour is strictly lower: cost[n] == cost - 1 (orthogonal) or cost - 2 (diagonal)
    DI = torch.tensor([1, 1, 0, -1, -1, -1, 0, 1, 0], device=dev)
    DJ = torch.tensor([0, 1, 1, 1, 0, -1, -1, -1, 0], device=dev)
    ii, jj = torch.meshgrid(torch.arange(G, device=dev), torch.arange(G, device=dev), indexing="ij")
    ni = (ii[None] + DI[d.long()]).clamp(0, G - 1)
    nj = (jj[None] + DJ[d.long()]).clamp(0, G - 1)
    nc = c[idx[:, None, None], ni, nj]
    drop = torch.where(d % 2 == 1, 2, 1)
    assert (nc[inner] == (c - drop)[inner]).all()
    occ_n, cost_n, flow_n = t2n(occ), t2n(cost), t2n(flow)
    for k in range(0, n, 97):
        ec, _, ef = oracle.flow_field(occ_n[k], int(goals[k, 0]), int(goals[k, 1]))
        assert np.array_equal(cost_n[k], ec) and np.array_equal(flow_n[k], ef), k


# ---------------------------------------------------------------------------------------------------
# step()/reset(): full env rollouts
# ---------------------------------------------------------------------------------------------------
def compare_env(env, orc, t, check_planes=False):
    obs = env._obs()
    assert np.array_equal(t2n(obs["local_map"]), orc.local_map), (t, "local_map")
    for name, got, exp in [("rel_goal", env.rel_goal, orc.rel_goal), ("velocity", env.velocity, orc.velocity),
                           ("pose", env.pose(), orc.pose), ("goal", env.goal(), orc.goal)]:
        g = t2n(got)
        assert np.allclose(g, exp, rtol=1e-5, atol=1e-6), (t, name)          # north-star tolerance
        assert np.array_equal(g.view(np.uint32), exp.view(np.uint32)), (t, name, "bit-exact")
    assert np.array_equal(t2n(env.episode()).astype(np.uint32), orc.episode), (t, "episode")
    assert np.array_equal(t2n(env.steps()), orc.steps), (t, "steps")
    if check_planes:
        assert np.array_equal(t2n(env.cost_field()), orc.cost), (t, "cost")
        assert np.array_equal(t2n(env.flow_image()), orc.flow), (t, "flow")
        assert np.array_equal(t2n(env.occupancy()), orc.occ), (t, "occ")
        assert np.array_equal(t2n(env.flow_dir()), orc.dir), (t, "dir")


def rollout_parity(ffmp, N, steps, seed=0, check_every=50, terminal_obs=True, **kw):
    env = ffmp.FFMPVectorEnv(N, seed=seed, terminal_obs=terminal_obs, **kw)
    okw = {k: v for k, v in kw.items() if k in ("grid", "window", "max_steps", "goal_mode", "p_occ", "block_shift", "env_id_base", "dt")}
    orc = oracle.OracleVectorEnv(N, seed=seed, **okw)
    env.reset()
    orc.reset()
    compare_env(env, orc, -1, check_planes=True)
    rng = np.random.default_rng(seed + 1)
    ndone = 0
    for t in range(steps):
        a = rng.integers(0, 28, N)
        _, reward, done, info = env.step(torch.as_tensor(a, device=env.device))
        orc.step(a)
        r = t2n(reward)
        assert np.allclose(r, orc.reward, rtol=1e-5, atol=1e-7), (t, "reward tol")
        assert np.array_equal(r.view(np.uint32), orc.reward.view(np.uint32)), (t, "reward")
        assert np.array_equal(t2n(done).astype(np.uint8), orc.done), (t, "done")
        assert np.array_equal(t2n(info["flags"]), orc.flags), (t, "flags")
        assert np.array_equal(t2n(info["terminal_relative_goal"]).view(np.uint32), orc.term_rel_goal.view(np.uint32)), t
        assert np.array_equal(t2n(info["terminal_velocity"]).view(np.uint32), orc.term_velocity.view(np.uint32)), t
        d = orc.done.astype(bool)
        if d.any():
            assert np.array_equal(t2n(info["episode_return"])[d].view(np.uint32), orc.fin_return[d].view(np.uint32)), t
            assert np.array_equal(t2n(info["episode_length"])[d], orc.fin_length[d]), t
            if terminal_obs:      # a15: the terminal observation of the envs that finished ([previous frame, terminal crop])
                assert np.array_equal(t2n(info["terminal_local_map"])[d], orc.term_local_map[d]), (t, "terminal_local_map")
        assert info["observe_t"] == env.config.dt
        ndone += int(d.sum())
        compare_env(env, orc, t, check_planes=(t % check_every == check_every - 1))
    assert env.error_word() == 0
    env.close()
    return ndone


def test_rollout_1k_steps_config3_shape(ffmp):
    """1k-step random-action rollout, 128x128 grids, W=100, goal re-sampled per reset (config 3 per-env shape)."""
    ndone = rollout_parity(ffmp, 48, 1000, seed=0, grid=128, window=100, check_every=100)
    assert ndone > 200            # many auto-resets (and background regenerations) were exercised


def test_rollout_config2_static_goal(ffmp):
    rollout_parity(ffmp, 64, 400, seed=3, grid=64, window=64, goal_mode=1, check_every=100)


def test_rollout_config1_default_grid(ffmp):
    """config 1: one env, the reference's 100x100 grid (ffmp.py:15), 1k random steps."""
    rollout_parity(ffmp, 1, 1000, seed=0, grid=100, window=100, check_every=250)


@pytest.mark.parametrize("ring,slots", [(2, 3), (3, 2), (5, 4), (8, 3)])
def test_rollout_ring_and_slot_variants(ffmp, ring, slots):
    # (the (3, 2) case also runs without the terminal-observation kernel: the default configuration of the env)
    rollout_parity(ffmp, 32, 150, seed=7, grid=64, window=32, ring=ring, slots=slots, max_steps=12, check_every=30)
    if (ring, slots) == (3, 2):
        rollout_parity(ffmp, 32, 60, seed=8, grid=64, window=32, ring=ring, slots=slots, max_steps=12, check_every=30, terminal_obs=False)


@pytest.mark.parametrize("slots,batch", [(12, 2), (8, 3), (4, 3), (16, 5), (5, 2)])
def test_rollout_grouped_regeneration(ffmp, cuda_device, slots, batch):
    """ffmp_cfg.regen_batch: the episode ends of `batch` consecutive ticks share one regeneration list and launch (and the
    ticks in between start as programmatic dependents of each other).  Results must not depend on the grouping: stepwise
    against the oracle with short episodes, then open-loop rollouts whose lengths leave groups unfinished at the joins."""
    kw = dict(grid=64, window=32, slots=slots, regen_batch=batch, max_steps=6)
    rollout_parity(ffmp, 32, 90, seed=17, check_every=31, terminal_obs=False, **kw)
    rollout_parity(ffmp, 24, 40, seed=18, check_every=13, p_occ=0.3, block_shift=0, **{**kw, "grid": 128, "window": 100})
    N = 48
    env = ffmp.FFMPVectorEnv(N, seed=19, **kw)
    orc = oracle.OracleVectorEnv(N, seed=19, grid=64, window=32, max_steps=6)
    env.reset(); orc.reset()
    rng = np.random.default_rng(3)
    done_total = 0
    for T in (1, 7, 40, 2 * batch, 2 * batch + 1, 33):
        acts = rng.integers(0, 28, (T, N))
        env.rollout(torch.as_tensor(acts, device=cuda_device))
        for t in range(T):
            orc.step(acts[t])
            done_total += int(orc.done.sum())
        compare_env(env, orc, T, check_planes=True)
    assert done_total > 8 * N and env.error_word() == 0
    env.close()


@pytest.mark.parametrize("slots,batch,term", [(16, 0, False), (8, 0, True), (6, 2, False)])
def test_rollout_graph_replay_matches_oracle(ffmp, cuda_device, slots, batch, term):
    """ffmp_rollout_graphed: T ticks and their background regenerations captured once per (buffer, T, ring / list phase) and
    replayed as one graph launch.  The action buffer is refilled in place between replays; phases cycle because T is not a
    multiple of the ring or list periods; plain steps and plain rollouts are interleaved with the replays."""
    N, T = 40, 10
    kw = dict(grid=64, window=32, slots=slots, regen_batch=batch, max_steps=7)
    env = ffmp.FFMPVectorEnv(N, seed=23, terminal_obs=term, **kw)
    orc = oracle.OracleVectorEnv(N, seed=23, grid=64, window=32, max_steps=7)
    env.reset(); orc.reset()
    rng = np.random.default_rng(9)
    buf = torch.zeros((T, N), dtype=torch.int64, device=cuda_device)
    launches0 = env.launch_count()
    for it in range(24):
        acts = rng.integers(0, 28, (T, N))
        buf.copy_(torch.as_tensor(acts, device=cuda_device))
        env.rollout(buf, graph=True)
        for t in range(T):
            orc.step(acts[t])
        compare_env(env, orc, it, check_planes=(it % 6 == 5))
        if term and orc.done.any():
            d = orc.done.astype(bool)
            assert np.array_equal(t2n(env._info()["terminal_local_map"])[d], orc.term_local_map[d]), (it, "terminal_local_map")
        if it % 5 == 4:           # plain calls in between shift the phases the next replays start from
            a1 = rng.integers(0, 28, (3, N))
            env.step(torch.as_tensor(a1[0], device=cuda_device)); orc.step(a1[0])
            env.rollout(torch.as_tensor(a1[1:], device=cuda_device))
            orc.step(a1[1]); orc.step(a1[2])
            compare_env(env, orc, it, check_planes=True)
    assert env.launch_count() - launches0 >= 24 * T and env.error_word() == 0
    env.close()


def test_rollout_with_regeneration_wait_forced(ffmp, monkeypatch):
    """FFMP_REGEN_WAIT=1: the tick always waits on the regeneration's event instead of trusting the completion word in
    mapped memory; same results either way."""
    monkeypatch.setenv("FFMP_REGEN_WAIT", "1")
    rollout_parity(ffmp, 32, 120, seed=27, grid=64, window=32, slots=3, max_steps=9, check_every=40)
    rollout_parity(ffmp, 16, 60, seed=28, grid=128, window=100, slots=12, p_occ=0.3, block_shift=0, check_every=30)


@pytest.mark.parametrize("grid,bs,p", [(128, 3, 0.1), (128, 0, 0.3), (116, 1, 0.1)])
def test_rollout_warp_kernel_regeneration(ffmp, monkeypatch, grid, bs, p):
    """Resets of a few envs and the background regeneration lists run on the four-warps-per-grid kernel
    (flow_field_quad_kernel); FFMP_FLOW_QUAD=0 keeps them on the warp-per-grid kernel that serves large batches.  Both
    must produce the oracle's planes (every other 96 < G <= 128 rollout test runs the quad kernel)."""
    monkeypatch.setenv("FFMP_FLOW_QUAD", "0")
    rollout_parity(ffmp, 24, 80, seed=31, grid=grid, window=100, p_occ=p, block_shift=bs, max_steps=20, check_every=20)
    monkeypatch.setenv("FFMP_FLOW_QUAD", "1")
    rollout_parity(ffmp, 24, 80, seed=31, grid=grid, window=100, p_occ=p, block_shift=bs, max_steps=20, check_every=20)


@pytest.mark.parametrize("dt,ring", [(0.1, 3), (0.3, 3), (0.3, 8), (0.16, 2)])
def test_rollout_long_steps_and_short_rings(ffmp, dt, ring):
    """Steps of up to 3.6 cells (dt = 0.3 s at 0.6 m/s) and rings that wrap every step / every other step (the older frame is
    refetched at the previous pose on a wrap).  Also the regression test of the superset-window experiment."""
    rollout_parity(ffmp, 32, 120, seed=41, grid=128, window=100, ring=ring, dt=dt, check_every=40)
    rollout_parity(ffmp, 24, 90, seed=42, grid=64, window=32, ring=ring, dt=dt, max_steps=15, check_every=30)


@pytest.mark.parametrize("switch", ["FFMP_STEP_FUSED=0", "FFMP_TICK_PDL=0", "FFMP_TRACE=1"])
def test_rollout_under_library_switches(ffmp, monkeypatch, switch):
    """Every environment switch that selects another launch path has a parity run: the two-kernel step (dynamics + observe, also
    what grids without a tensor map use), step kernels without the programmatic-dependent attribute, the tracing variant of the
    step kernel (FFMP_FLOW_QUAD, FFMP_ACT_PARAM, FFMP_HOST_IO, FFMP_REGEN_WAIT, FFMP_FLOW_ROWS, FFMP_SCAN_GENERIC: tests above)."""
    k, v = switch.split("=")
    monkeypatch.setenv(k, v)
    rollout_parity(ffmp, 24, 60, seed=61, grid=128, window=100, max_steps=20, check_every=30)
    rollout_parity(ffmp, 16, 40, seed=62, grid=64, window=32, ring=3, slots=3, max_steps=9, check_every=20)


def test_rollout_dense_obstacles_short_episodes(ffmp):
    """p=0.3 per-cell noise: episodes of a few steps, so nearly every step regenerates slots."""
    rollout_parity(ffmp, 32, 200, seed=9, grid=128, window=100, p_occ=0.3, block_shift=0, check_every=50)


@pytest.mark.parametrize("grid,bs,p", [(112, 1, 0.05), (124, 2, 0.15), (128, 5, 0.2), (104, 7, 0.1)])
def test_rollout_interleaved_generator_block_sizes(ffmp, grid, bs, p):
    """Grids in (96, 128] use the column-interleaved layout; its in-kernel scenario generator has separate code for
    block_shift < 2, 2..6 and >= 7, and the row stores differ for G % 16 != 0."""
    rollout_parity(ffmp, 24, 120, seed=13, grid=grid, window=32, p_occ=p, block_shift=bs, max_steps=15, check_every=40)


def test_sharded_env_ids(ffmp):
    """Two shards (env_id_base 0 and 8) reproduce one batch of 16: results depend on global ids only."""
    rollout_parity(ffmp, 8, 60, seed=21, grid=64, window=32, env_id_base=8, check_every=30)


def test_masked_reset(ffmp, cuda_device):
    N = 16
    env = ffmp.FFMPVectorEnv(N, seed=4, grid=64, window=32)
    orc = oracle.OracleVectorEnv(N, seed=4, grid=64, window=32)
    env.reset(); orc.reset()
    rng = np.random.default_rng(0)
    for t in range(40):
        a = rng.integers(0, 28, N)
        env.step(torch.as_tensor(a, device=cuda_device)); orc.step(a)
        if t % 7 == 3:
            m = rng.random(N) < 0.4
            env.reset(torch.as_tensor(m, device=cuda_device)); orc.reset(m)
        compare_env(env, orc, t, check_planes=(t % 10 == 9))
    # a second full reset restarts from episode 0 with identical results
    first = {k: t2n(v).copy() for k, v in env.reset().items()}
    orc.reset()
    compare_env(env, orc, 100, check_planes=True)
    again = {k: t2n(v).copy() for k, v in env.reset().items()}
    for k in first:
        assert np.array_equal(first[k], again[k])
    env.close()


def test_rollout_api_and_host_step(ffmp, cuda_device):
    N, T = 32, 25
    env = ffmp.FFMPVectorEnv(N, seed=12, grid=64, window=32)
    orc = oracle.OracleVectorEnv(N, seed=12, grid=64, window=32)
    env.reset(); orc.reset()
    rng = np.random.default_rng(5)
    acts = rng.integers(0, 28, (T, N))
    env.rollout(torch.as_tensor(acts, device=cuda_device))
    for t in range(T):
        orc.step(acts[t])
    compare_env(env, orc, T)
    host_a = torch.as_tensor(rng.integers(0, 28, N)).pin_memory()
    obs, reward, done, info = env.step_host(host_a)
    orc.step(host_a.numpy())
    assert reward.device.type == "cpu" and np.array_equal(reward.numpy().view(np.uint32), orc.reward.view(np.uint32))
    assert np.array_equal(done.numpy().astype(np.uint8), orc.done)
    assert np.array_equal(obs["relative_goal"].numpy().view(np.uint32), orc.rel_goal.view(np.uint32))
    assert np.array_equal(t2n(obs["local_map"]), orc.local_map)
    env.close()


@pytest.mark.parametrize("host_io,pinned,act_param", [(0, True, 1), (1, True, 1), (2, True, 1), (2, False, 1), (0, True, 0),
                                                       (1, True, 0), (1, False, 0)])
def test_host_step_paths(ffmp, cuda_device, monkeypatch, host_io, pinned, act_param):
    """ffmp_step_host through the copy engines (FFMP_HOST_IO=0), with the results written to pinned host memory by the
    export kernel (1), with the actions read in place as well (2), and with pageable buffers (falls back to copies); the
    actions of up to 4096 envs ride in the step kernel's launch as a by-value parameter unless FFMP_ACT_PARAM=0 (then they
    are copied to the device first).  200 steps with auto-resets, every host-side result bit-compared with the oracle;
    step_async / step_wait split."""
    monkeypatch.setenv("FFMP_HOST_IO", str(host_io))
    monkeypatch.setenv("FFMP_ACT_PARAM", str(act_param))
    N = 64
    env = ffmp.FFMPVectorEnv(N, seed=31, grid=64, window=32, max_steps=20)
    orc = oracle.OracleVectorEnv(N, seed=31, grid=64, window=32, max_steps=20)
    if not pinned:      # pageable result block with the packed layout
        block = torch.zeros((22 * N + 16,), dtype=torch.uint8)
        r, g, v, dn, fl = env._split_out_block(block, N)
        env._host = {"reward": r, "rel_goal": g, "velocity": v, "done": dn, "flags": fl, "done_bool": dn.view(torch.bool), "info": {"flags": fl}}
        import ctypes as C
        env._host_ptrs = tuple(C.c_void_p(env._host[k].data_ptr()) for k in ("reward", "done", "flags", "rel_goal", "velocity"))
    env.reset(); orc.reset()
    rng = np.random.default_rng(8)
    bufs = [torch.zeros(N, dtype=torch.int64) for _ in range(3)]
    if pinned:
        bufs = [b.pin_memory() for b in bufs]
    ndone = 0
    for t in range(200):
        a = bufs[t % 3]
        a.copy_(torch.as_tensor(rng.integers(0, 28, N)))
        if t % 2:
            obs, reward, done, info = env.step_host(a)
        else:
            env.step_async(a)
            obs, reward, done, info = env.step_wait()
        orc.step(a.numpy())
        assert reward.device.type == "cpu"
        assert np.array_equal(reward.numpy().view(np.uint32), orc.reward.view(np.uint32)), t
        assert np.array_equal(done.numpy().astype(np.uint8), orc.done), t
        assert np.array_equal(info["flags"].numpy(), orc.flags), t
        assert np.array_equal(obs["relative_goal"].numpy().view(np.uint32), orc.rel_goal.view(np.uint32)), t
        assert np.array_equal(obs["velocity"].numpy().view(np.uint32), orc.velocity.view(np.uint32)), t
        ndone += int(orc.done.sum())
        if t % 20 == 19:
            compare_env(env, orc, t)
    assert ndone > 100 and env.error_word() == 0
    # a second wait without an async is a no-op; two asyncs without a wait are refused
    assert env._L.ffmp_step_host_wait(env._h) == 0
    env.step_async(bufs[0])
    with pytest.raises(ffmp.native.NativeError):
        env.step_async(bufs[1])
    env.step_wait()
    env.close()


def test_invalid_actions_and_errors(ffmp, cuda_device):
    env = ffmp.FFMPVectorEnv(4, seed=1, grid=64, window=32)
    with pytest.raises(ffmp.native.NativeError):
        env.step(torch.zeros(4, dtype=torch.int64, device=cuda_device))       # step before reset
    env.reset()
    env.step(torch.tensor([3, 99, -5, 27], device=cuda_device))
    assert env.error_word() & 1
    # the same through the host-buffer step (actions inside the launch, one byte per env: out of range must stay out of range)
    env2 = ffmp.FFMPVectorEnv(4, grid=64, window=32)
    env2.reset()
    env2.step_host(torch.tensor([3, 0, 27, 5]).pin_memory())
    assert env2.error_word() == 0
    env2.step_host(torch.tensor([3, 255 + 28, -(2 ** 40), 2 ** 33 + 1]).pin_memory())
    assert env2.error_word() & 1
    env2.close()
    with pytest.raises(ValueError):
        env.step(torch.zeros(5, dtype=torch.int64, device=cuda_device))
    # ffmp_bind rejects output pointers the kernels' float2 / 4-byte stores cannot take
    import ctypes as C
    names = ("cost", "flow", "scen", "state", "frames", "rel_goal", "velocity", "reward", "done", "flags", "term_rel_goal",
             "term_velocity", "fin_return", "fin_length")
    ptrs = {n: getattr(env, n).data_ptr() for n in names}
    L = ffmp.native.lib()
    for bad, off in (("rel_goal", 4), ("term_velocity", 4), ("reward", 2), ("flow", 8)):
        b = ffmp.native.Buffers(**{**ptrs, bad: ptrs[bad] + off}, workspace=env._workspace.data_ptr())
        assert L.ffmp_bind(env._h, C.byref(b)) < 0 and b"aligned" in L.ffmp_last_error(), bad
    b = ffmp.native.Buffers(**ptrs, workspace=env._workspace.data_ptr())
    assert L.ffmp_bind(env._h, C.byref(b)) == 0
    env.close()
    with pytest.raises(ffmp.native.NativeError):
        ffmp.FFMPVectorEnv(4, grid=1024)                                        # unsupported in this build
    with pytest.raises(ffmp.native.NativeError):
        ffmp.FFMPVectorEnv(4, window=50)


# ---------------------------------------------------------------------------------------------------
# reference surface: FFMP compat object against the reference's golden answers
# ---------------------------------------------------------------------------------------------------
def test_compat_object_against_reference_golden(ffmp, golden):
    env = ffmp.make("FFMP-v0")
    for case in golden["is_collision"]:
        m = np.zeros((100, 100), np.int32)
        for (i, j) in case["cells"]:
            m[i, j] = case["value"]
        assert env.is_collision(m) == case["expect"], case
    for case in golden["is_collision2"]:
        assert env.is_collision2(case["scan"]) == case["expect"], case
    for case in golden["is_goal"]:
        if abs(case["d"] - 0.5) > 1e-6:                       # fp32 cannot resolve 0.5 +- 1e-7
            assert env.is_goal(case["d"]) == case["expect"], case
    assert env.is_goal(0.5) is False and env.is_goal(0.49) is True
    for case in golden["is_done"]:
        assert env.is_done(case["col"], case["goal"]) == case["expect"]
    for seq in golden["reward_sequences"]:
        for s in seq:
            got = env.reward_calculator([s["d"], 0.0], s["col"], s["goal"], s["is_first"])
            assert abs(got - s["expect"]) <= 1e-5 * max(1.0, abs(s["expect"])), s
    for case in golden["rewarder"]:
        m = np.zeros((100, 100), np.int32)
        for (i, j) in case["cells"]:
            m[i, j] = 255
        r, d = env.rewarder(m, case["rel_goal"], case["is_first"])
        assert abs(r - case["expect"][0]) <= 1e-5 and d == case["expect"][1], case
    for case in golden["rewarder2"]:
        r, d, g = env.rewarder2(case["scan"], case["rel_goal"], case["is_first"])
        assert abs(r - case["expect"][0]) <= 1e-5 and [d, g] == case["expect"][1:], case


def test_gym_make_through_the_gym_ffmp_shim(ffmp, golden, monkeypatch):
    """train.py:35,456: `import gym_ffmp; gym.make('FFMP-v0')` — with a `gym` importable (here the stand-in registry of
    oracle/gym_shim, test infrastructure) the id resolves to the CUDA-backed FFMP object; train.py:577's only env call,
    rewarder2, answers as the reference does."""
    import importlib
    import sys
    monkeypatch.syspath_prepend(os.path.join(ROOT, "oracle", "gym_shim"))
    monkeypatch.syspath_prepend(ROOT)
    for m in [k for k in sys.modules if k == "gym" or k.startswith("gym.") or k == "gym_ffmp" or k.startswith("gym_ffmp.")]:
        monkeypatch.delitem(sys.modules, m)
    gym = importlib.import_module("gym")
    gym_ffmp = importlib.import_module("gym_ffmp")
    assert "gym" in gym_ffmp.REGISTERED_WITH
    env = gym.make("FFMP-v0")
    assert isinstance(env, ffmp.FFMP)
    for case in golden["rewarder2"]:
        r, d, g = env.rewarder2(case["scan"], case["rel_goal"], case["is_first"])
        assert abs(r - case["expect"][0]) <= 1e-5 and [d, g] == case["expect"][1:], case
    obs = env.reset()
    assert obs["local_map"].shape == (100, 100, 1) and obs["local_map"].dtype == np.int32
    env.close()


def test_env_collision_equals_reference_is_collision_on_crop(ffmp, cuda_device):
    """The env's footprint test on the global grid == FFMP.is_collision on the ego-centred occupancy crop
    (reference footprint pinned by tests/golden)."""
    N = 64
    env = ffmp.FFMPVectorEnv(N, seed=33, grid=128, window=100, ring=2)
    env.reset()
    rng = np.random.default_rng(1)
    L = oracle.lib()
    import ctypes as C
    checked = 0
    for t in range(60):
        a = rng.integers(7, 28, N)
        pose_before = t2n(env.pose()).copy()
        occ_before = t2n(env.occupancy()).copy()
        _, _, done, info = env.step(torch.as_tensor(a, device=cuda_device))
        flags = t2n(info["flags"])
        # rebuild the occupancy crop at the pose the step evaluated (terminal pose for finished envs)
        for e in range(N):
            v, w = [(c.linear_v, c.angular_v) for c in [env.action.commander(int(a[e]))]][0]
            x, y, yaw = pose_before[e]
            s, c = oracle.sincos(yaw)
            nx = np.float32(x + np.float32(np.float32(np.float32(v) * c) * np.float32(0.1)))
            ny = np.float32(y + np.float32(np.float32(np.float32(v) * s) * np.float32(0.1)))
            ci = int(np.floor(np.float32(nx * np.float32(20.0)) + np.float32(0.5)))
            cj = int(np.floor(np.float32(ny * np.float32(20.0)) + np.float32(0.5)))
            crop = oracle.crop(np.where(occ_before[e] != 0, 255, 0).astype(np.uint8), 100, ci, cj).astype(np.int32)
            ref = L.orc_ref_is_collision(crop.ctypes.data_as(C.POINTER(C.c_int32)), 100, 0.05, 5.0, 0.13)
            assert bool(ref) == bool(flags[e] & 1), (t, e)
            checked += 1
    assert checked == 60 * N
    env.close()


# ---------------------------------------------------------------------------------------------------
# kernel variants behind environment switches (read by the library at create / launch time)
# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("G", [32, 64, 128])
def test_flow_field_rows_kernel_on_small_grids(ffmp, cuda_device, G, monkeypatch):
    """FFMP_FLOW_ROWS=1: the CTA-per-grid / thread-per-row kernel of the large-map path, on grids the warp kernel normally
    takes; with and without the cost output."""
    monkeypatch.setenv("FFMP_FLOW_ROWS", "1")
    cases = special_cases(G)
    cost, flow = run_flow(ffmp, cuda_device, [c[1] for c in cases], [c[2] for c in cases])
    for k, (name, occ, goal) in enumerate(cases):
        ec, ed, ef = oracle.flow_field(occ, goal[0], goal[1])
        assert np.array_equal(cost[k], ec), (G, name, "cost", int((cost[k] != ec).sum()))
        assert np.array_equal(flow[k], ef), (G, name, "flow", int((flow[k] != ef).sum()))
    _, flow2 = run_flow(ffmp, cuda_device, [c[1] for c in cases], [c[2] for c in cases], want_cost=False)
    assert np.array_equal(flow2, flow)


def test_rollout_rows_kernel_regeneration(ffmp, monkeypatch):
    """FFMP_FLOW_ROWS=1 inside the env: reset and background regeneration through the row kernel."""
    monkeypatch.setenv("FFMP_FLOW_ROWS", "1")
    rollout_parity(ffmp, 48, 300, seed=11, grid=128, window=100, max_steps=20, check_every=50)


def test_flow_field_large_without_cost(ffmp, cuda_device):
    occ, _, _, cells = oracle.scenario(5, 0, 0, 256, p_occ=0.2, block_shift=2)
    _, flow = run_flow(ffmp, cuda_device, [occ], [(cells[2], cells[3])], want_cost=False)
    assert np.array_equal(flow[0], oracle.flow_field(occ, cells[2], cells[3])[2])


def test_learner_input_matches_float_cast(ffmp, cuda_device):
    """ffmp_learner_input == the reference's observe_m: cat of the last two frames, `.float()` (train.py:474-486, 539-545)."""
    env = ffmp.FFMPVectorEnv(37, grid=64, window=32, ring=4, seed=2)
    obs = env.reset()
    rng = np.random.default_rng(2)
    for t in range(9):      # crosses two ring wraps
        obs, _, _, _ = env.step(torch.as_tensor(rng.integers(0, 28, 37), device=cuda_device))
        want = obs["local_map"].float()
        assert torch.equal(env.learner_input(), want)
        assert torch.equal(env.learner_input(dtype=torch.bfloat16).float(), want)
        assert torch.allclose(env.learner_input(scale=1.0 / 255.0), want / 255.0, rtol=0, atol=1e-7)
    env.close()


# ---------------------------------------------------------------------------------------------------
# L: LiDAR scan synthesis (SPEC.md §9)
# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("G,beams,rmax,flow_mode", [(128, 360, 3.5, True), (64, 7, 1.0, False), (100, 1000, 10.0, True), (32, 1, 3.5, False)])
def test_scan_operator_bit_exact(ffmp, cuda_device, G, beams, rmax, flow_mode):
    rng = np.random.default_rng(G + beams)
    n = 24
    maps, poses = [], []
    for k in range(n):
        occ, start, goal, cells = oracle.scenario(5, k, 1, G, p_occ=0.15)
        if flow_mode:
            _, _, m = oracle.flow_field(occ, cells[2], cells[3])
        else:
            m = occ * np.uint8(rng.integers(1, 255))          # any non-zero value is occupied
        maps.append(m)
        if k % 6 == 5:      # anywhere, including occupied cells, the border and outside the grid
            poses.append([rng.uniform(-0.3, G * 0.05 + 0.3), rng.uniform(-0.3, G * 0.05 + 0.3), rng.uniform(-3.14, 3.14)])
        else:               # a free cell, off-centre
            free = np.argwhere(occ == 0)
            i, j = free[rng.integers(0, len(free))]
            poses.append([i * 0.05 + rng.uniform(-0.024, 0.024), j * 0.05 + rng.uniform(-0.024, 0.024), rng.uniform(-3.14, 3.14)])
    poses[0][2] = 0.0                                          # axis-aligned beams (a direction cosine of exactly 0 / 1)
    poses[1] = [float(np.float32(5 * 0.05)), float(np.float32(7 * 0.05)), float(np.float32(np.pi / 2))]
    maps_n, poses_n = np.stack(maps), np.array(poses, np.float32)
    scan, hit = ffmp.ops.scan(torch.as_tensor(maps_n, device=cuda_device), torch.as_tensor(poses_n, device=cuda_device),
                              beams=beams, range_max=rmax, flow_mode=flow_mode)
    scan, hit = t2n(scan), t2n(hit)
    for k in range(n):
        er, eh = oracle.scan(maps_n[k], poses_n[k], beams, rmax, flow_mode)
        assert np.array_equal(scan[k].view(np.uint32), er.view(np.uint32)), k
        assert hit[k] == eh, k
    assert (scan > 0).any()


def test_scan_kernel_variants_agree(ffmp, cuda_device, monkeypatch):
    """Default = shared-memory window kernel; FFMP_SCAN_GENERIC=1 = global-load kernel (used for range_max beyond the window)."""
    env = ffmp.FFMPVectorEnv(16, seed=2, grid=128, window=100)
    env.reset()
    for beams in (360, 1000):
        a, ha = env.scan(beams, 3.5)
        monkeypatch.setenv("FFMP_SCAN_GENERIC", "1")
        c, hc = env.scan(beams, 3.5)
        monkeypatch.delenv("FFMP_SCAN_GENERIC")
        assert torch.equal(a.view(torch.int32), c.view(torch.int32)) and torch.equal(ha, hc), beams
    env.close()


def test_scan_env_rollout_and_rewarder2(ffmp, cuda_device):
    """env.scan() at every step of a rollout equals the oracle's scan of the oracle env's pose and flow image, and feeding
    it to the batched FFMP.rewarder2 gives the collision flag the reference function computes from the same list."""
    N = 32
    env = ffmp.FFMPVectorEnv(N, seed=17, grid=128, window=100)
    orc = oracle.OracleVectorEnv(N, seed=17, grid=128, window=100)
    env.reset(); orc.reset()
    rng = np.random.default_rng(3)
    hits = 0
    for t in range(60):
        a = rng.integers(0, 28, N)
        env.step(torch.as_tensor(a, device=cuda_device)); orc.step(a)
        scan, hit = env.scan(beams=360, range_max=3.5)
        es, eh = orc.scan(360, 3.5)
        assert np.array_equal(t2n(scan).view(np.uint32), es.view(np.uint32)), t
        assert np.array_equal(t2n(hit), eh), t
        hits += int(eh.sum())
        if t % 20 == 0:
            # rewarder2 takes f64 ranges with NaN = None: drop inf / 0 as the trainer's callback does (train.py:144-150)
            s64 = scan.double()
            s64 = torch.where(torch.isinf(s64) | (s64 == 0), torch.full_like(s64, float("nan")), s64)
            d_first = torch.zeros(N, dtype=torch.float32, device=cuda_device)
            reward, done, flags = ffmp.ops.rewarder2(s64, env.rel_goal, torch.ones(N, dtype=torch.uint8, device=cuda_device), d_first)
            assert np.array_equal(t2n(flags) & 1, eh), t
    env.close()


# ---------------------------------------------------------------------------------------------------
# learner feed through the peer-memory push kernel (world size 1 here: the destination is the rank's own buffer; the
# cross-GPU path is exercised by tools/feed_bench.py on a multi-GPU box)
# ---------------------------------------------------------------------------------------------------
def test_learner_feed_push_matches_pack_transitions(ffmp, cuda_device):
    from flow_field_based_motion_planner_b200 import sharding
    N, W = 48, 32
    env = ffmp.FFMPVectorEnv(N, seed=5, grid=64, window=W, ring=4, max_steps=15)
    env.reset()
    feed = sharding.LearnerFeed(env)
    assert feed.world == 1 and feed.block == sharding.transition_nbytes(N, W)
    rng = np.random.default_rng(1)
    for t in range(40):                       # ring wraps, auto-resets, both buffers and the credit path (seq > 2)
        obs, reward, done, _ = env.step(torch.as_tensor(rng.integers(0, 28, N), device=cuda_device))
        feed.push()
        g_obs, g_reward, g_done = feed.wait()
        expect = sharding.pack_transitions(obs, reward, done)
        e_obs, e_reward, e_done = sharding.unpack_transitions(expect, N, W)
        assert torch.equal(g_obs["local_map"][0], e_obs["local_map"]), t
        assert torch.equal(g_obs["relative_goal"][0].view(torch.int32), e_obs["relative_goal"].view(torch.int32)), t
        assert torch.equal(g_obs["velocity"][0].view(torch.int32), e_obs["velocity"].view(torch.int32)), t
        assert torch.equal(g_reward[0].view(torch.int32), e_reward.view(torch.int32)), t
        assert torch.equal(g_done[0] != 0, e_done), t
        # the raw slot is the packed block, byte for byte
        raw = feed._mem[(feed.seq & 1) * feed.buffer_stride:][:feed.block]
        assert torch.equal(raw, expect), t
        feed.release()
    assert feed.error_word() == 0
    feed.close()
    env.close()


def test_learner_feed_credit_timeout_sets_error_bit(ffmp, cuda_device):
    """A producer that is two pushes ahead of an unreleased buffer waits for the credit; the bounded spin gives up and flags it."""
    from flow_field_based_motion_planner_b200 import sharding
    env = ffmp.FFMPVectorEnv(8, seed=5, grid=64, window=32)
    env.reset()
    feed = sharding.LearnerFeed(env, timeout_s=0.05)
    for _ in range(3):                        # the third push needs buffer 1 back, nobody released it
        feed.push()
    assert feed.error_word() & 2
    feed.close()
    env.close()


@pytest.mark.parametrize("planes", [True, False])
def test_env_checkpoint_resume_is_bit_identical(ffmp, cuda_device, planes):
    """state_dict() in the middle of a rollout, then the same actions on (a) the original env and (b) a fresh env that
    loaded the checkpoint: identical results step by step, with and without the flow / cost planes in the checkpoint."""
    N = 32
    kw = dict(seed=11, grid=64, window=32, ring=4, slots=3, max_steps=12)
    env = ffmp.FFMPVectorEnv(N, **kw)
    env.reset()
    rng = np.random.default_rng(2)
    for _ in range(37):
        env.step(torch.as_tensor(rng.integers(0, 28, N), device=cuda_device))
    sd = env.state_dict(planes=planes)
    other = ffmp.FFMPVectorEnv(N, **kw)
    other.load_state_dict(sd)
    for t in range(60):
        a = torch.as_tensor(rng.integers(0, 28, N), device=cuda_device)
        o1, r1, d1, i1 = env.step(a)
        o2, r2, d2, i2 = other.step(a)
        assert torch.equal(r1.view(torch.int32), r2.view(torch.int32)) and torch.equal(d1, d2), t
        assert torch.equal(o1["local_map"], o2["local_map"]), t
        assert torch.equal(o1["relative_goal"].view(torch.int32), o2["relative_goal"].view(torch.int32)), t
        assert torch.equal(i1["flags"], i2["flags"]), t
    assert torch.equal(env.cost_field(), other.cost_field()) and torch.equal(env.flow_image(), other.flow_image())
    env.close(); other.close()


def test_replay_ring_stores_and_samples_transitions(ffmp, cuda_device):
    """Device replay ring (ReplayMemory / Transition, train.py:44,212-228): every stored field equals what the env returned
    at that step, FIFO overwrite at capacity, samples only pair consecutive pushes that are both still resident."""
    N, W, T = 16, 32, 6
    env = ffmp.FFMPVectorEnv(N, seed=23, grid=64, window=W, ring=3, max_steps=9)
    ring = ffmp.ReplayRing(env, T)
    obs = env.reset()
    ring.push()
    log = [dict(m=obs["local_map"].clone(), g=obs["relative_goal"].clone(), v=obs["velocity"].clone(), a=None, r=None, d=None)]
    rng = np.random.default_rng(6)
    gen = torch.Generator(device=cuda_device); gen.manual_seed(1)
    for t in range(1, 30):
        a = torch.as_tensor(rng.integers(0, 28, N), device=cuda_device)
        obs, reward, done, _ = env.step(a)
        ring.push(a)
        log.append(dict(m=obs["local_map"].clone(), g=obs["relative_goal"].clone(), v=obs["velocity"].clone(), a=a.clone(),
                        r=reward.clone(), d=done.clone()))
        assert len(ring) == (min(t + 1, T) - 1) * N
        batch = ring.sample(64, generator=gen)
        k, e = batch["index"][:, 0].tolist(), batch["index"][:, 1].tolist()
        assert min(k) >= max(1, t - T + 2) and max(k) <= t
        for b in range(64):
            s0, s1 = log[k[b] - 1], log[k[b]]
            assert torch.equal(batch["state_m"][b], s0["m"][e[b]]) and torch.equal(batch["observe_m"][b], s1["m"][e[b]])
            assert torch.equal(batch["state_g"][b], s0["g"][e[b]]) and torch.equal(batch["observe_g"][b], s1["g"][e[b]])
            assert torch.equal(batch["state_v"][b], s0["v"][e[b]]) and torch.equal(batch["observe_v"][b], s1["v"][e[b]])
            assert batch["action"][b] == s1["a"][e[b]] and batch["reward"][b] == s1["r"][e[b]]
            assert bool(batch["done"][b]) == bool(s1["d"][e[b]])
    env.close()


def test_rollout_full_size_properties(ffmp, cuda_device):
    """BASELINE configs[2] at its full per-GPU size (4096 envs x 128 x 128, W = 100): size-independent properties instead of the
    oracle — one batch of 4096 equals two shards of 2048 (global env ids), and every step's outputs are self-consistent."""
    N, T = 4096, 120
    kw = dict(grid=128, window=100, seed=77)
    full = ffmp.FFMPVectorEnv(N, **kw)
    lo = ffmp.FFMPVectorEnv(N // 2, env_id_base=0, **kw)
    hi = ffmp.FFMPVectorEnv(N // 2, env_id_base=N // 2, **kw)
    for env in (full, lo, hi):
        env.reset()
    # a strided sample of the 4096 envs against the oracle, bit for bit, at the full size (every random draw is keyed by the
    # global env id, so a one-env oracle with env_id_base = id replays env `id` of the batch)
    sample = [0, 1, 511, 1024, 2047, 2048, 3333, 4095]
    orcs = [oracle.OracleVectorEnv(1, grid=128, window=100, seed=77, env_id_base=i) for i in sample]
    for o in orcs:
        o.reset()
    sidx = torch.as_tensor(sample, device=cuda_device)
    gen = torch.Generator(device=cuda_device); gen.manual_seed(5)
    ends = 0
    for t in range(T):
        a = torch.randint(0, 28, (N,), generator=gen, device=cuda_device)
        obs, reward, done, info = full.step(a)
        a_s = t2n(a[sidx])
        for k, o in enumerate(orcs):
            o.step(a_s[k:k + 1])
        assert np.array_equal(t2n(reward[sidx]).view(np.uint32), np.concatenate([o.reward for o in orcs]).view(np.uint32)), (t, "sample reward")
        assert np.array_equal(t2n(done[sidx]).astype(np.uint8), np.concatenate([o.done for o in orcs])), (t, "sample done")
        assert np.array_equal(t2n(obs["relative_goal"][sidx]).view(np.uint32), np.concatenate([o.rel_goal for o in orcs]).view(np.uint32)), (t, "sample goal")
        if t % 10 == 9:
            assert np.array_equal(t2n(obs["local_map"][sidx]), np.concatenate([o.local_map for o in orcs])), (t, "sample local_map")
        o1, r1, d1, i1 = lo.step(a[:N // 2].contiguous())
        o2, r2, d2, i2 = hi.step(a[N // 2:].contiguous())
        assert torch.equal(reward.view(torch.int32), torch.cat([r1, r2]).view(torch.int32)), t
        assert torch.equal(done, torch.cat([d1, d2])), t
        assert torch.equal(info["flags"], torch.cat([i1["flags"], i2["flags"]])), t
        assert torch.equal(obs["relative_goal"].view(torch.int32), torch.cat([o1["relative_goal"], o2["relative_goal"]]).view(torch.int32)), t
        flags = info["flags"]
        assert torch.equal(done, flags != 0), t                                   # done = collision | goal | truncated
        col, goal = (flags & 1) != 0, (flags & 2) != 0
        # reward = (goal ? 1 : 0.05 (d_first - d)) + (col ? -1 : 0) - 0.05 (ffmp.py:130-157): bounded by the map diagonal
        assert bool(((reward >= -1.05 - 0.05 * 10.0) & (reward <= 0.95 + 1e-6)).all()), t
        assert bool((reward[goal & ~col] == 0.95).all()) and bool((reward[col & ~goal] <= -1.05 + 0.05 * 10.0).all()), t
        assert bool((info["terminal_relative_goal"][:, 0][goal] < 0.5).all()), t   # is_goal: dist < 0.5 (ffmp.py:120-127)
        assert bool((full.steps() <= 200).all()) and bool((full.steps()[done] == 0).all()), t
        assert bool((obs["velocity"][done] == 0).all()), t                         # first tick of an episode (train.py:183-184)
        lm = obs["local_map"]
        assert bool((lm[:, 1, 50, 50] != 255).all()), t                            # the robot never stands on an occupied cell
        assert bool((lm[done][:, 0] == lm[done][:, 1]).all()), t                   # both frames equal on the first tick (train.py:475-478)
        ends += int(done.sum())
        if t % 40 == 39:
            assert torch.equal(obs["local_map"][:N // 2], o1["local_map"]) and torch.equal(obs["local_map"][N // 2:], o2["local_map"]), t
    assert ends > 5000 and full.error_word() == 0
    # the integration field of every current scenario: 0 at the goal cell, INF exactly on occupied cells of reached components
    cost, occ = full.cost_field(), full.occupancy()
    gc = full.scen.view(torch.int32)[full.episode().long() % full.config.slots, torch.arange(N, device=cuda_device)][:, 5:7].long()
    assert bool((cost[torch.arange(N, device=cuda_device), gc[:, 0], gc[:, 1]] == 0).all())
    assert bool((cost[occ != 0] == INF).all())
    for env in (full, lo, hi):
        env.close()


def test_single_env_gym_api_matches_oracle(ffmp):
    """gym.make('FFMP-v0')-style object: reset() / step(action id) in the old 4-tuple API on the reference's 100 x 100 map
    (BASELINE config 1), numpy observations with the declared shapes / dtypes, equal to the oracle step by step."""
    env = ffmp.make("FFMP-v0", seed=7)
    orc = oracle.OracleVectorEnv(1, grid=100, window=100, seed=7)
    obs = env.reset(); orc.reset()
    assert obs["local_map"].shape == (100, 100, 1) and obs["local_map"].dtype == np.int32
    assert env.observation_space.spaces["local_map"].shape == (100, 100, 1)
    assert np.array_equal(obs["local_map"][:, :, 0], orc.local_map[0, 1])
    rng = np.random.default_rng(3)
    ends = 0
    for t in range(300):
        a = int(rng.integers(0, 28))
        obs, reward, done, info = env.step(a)
        orc.step(np.array([a]))
        assert np.float32(reward).view(np.uint32) == orc.reward.view(np.uint32)[0], t
        assert done == bool(orc.done[0]) and info["is_collision"] == bool(orc.flags[0] & 1) and info["is_goal"] == bool(orc.flags[0] & 2), t
        assert np.array_equal(obs["local_map"][:, :, 0], orc.local_map[0, 1]) and np.array_equal(info["local_map_stack"], orc.local_map[0]), t
        assert np.array_equal(obs["relative_goal"].view(np.uint32), orc.rel_goal[0].view(np.uint32)), t
        assert np.array_equal(obs["velocity"].view(np.uint32), orc.velocity[0].view(np.uint32)), t
        ends += int(done)
    assert ends >= 3
    env.close()
