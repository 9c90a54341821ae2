"""CPU, world_size 2, gloo: the multi-GPU host logic (shard ranges, learner-feed all-gather, episode-stat
all-reduce).  The shards are driven by the oracle here (no GPU in this container); on GPUs the same code
moves the CUDA env's tensors over NCCL."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from flow_field_based_motion_planner_b200 import sharding


def test_shard_range():
    assert [sharding.shard_range(10, r, 4) for r in range(4)] == [(0, 3), (3, 6), (6, 8), (8, 10)]
    assert [sharding.shard_range(8192, r, 2) for r in range(2)] == [(0, 4096), (4096, 8192)]
    for world in (1, 2, 3, 8):
        spans = [sharding.shard_range(513, r, world) for r in range(world)]
        assert spans[0][0] == 0 and spans[-1][1] == 513
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))


def test_pack_unpack_roundtrip():
    n, w = 5, 8
    g = torch.Generator().manual_seed(0)
    obs = {"local_map": torch.randint(0, 256, (n, 2, w, w), generator=g, dtype=torch.uint8),
           "relative_goal": torch.randn(n, 2, generator=g), "velocity": torch.randn(n, 2, generator=g)}
    reward, done = torch.randn(n, generator=g), torch.rand(n, generator=g) < 0.5
    buf = sharding.pack_transitions(obs, reward, done)
    assert buf.numel() == sharding.transition_nbytes(n, w)
    o2, r2, d2 = sharding.unpack_transitions(buf, n, w)
    assert all(torch.equal(o2[k], obs[k]) for k in obs) and torch.equal(r2, reward) and torch.equal(d2, done)
    # a strided ring view packs identically
    ring = torch.zeros((n, 6, w, w), dtype=torch.uint8)
    ring[:, 2:4] = obs["local_map"]
    obs_view = dict(obs, local_map=ring.narrow(1, 2, 2))
    assert torch.equal(sharding.pack_transitions(obs_view, reward, done), buf)


def _worker(rank, world, port, q):
    import oracle
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        total, W = 8, 16
        lo, hi = sharding.shard_range(total, rank, world)
        env = oracle.OracleVectorEnv(hi - lo, grid=32, window=W, seed=5, env_id_base=lo, max_steps=6, block_shift=2)
        env.reset()
        stats = sharding.EpisodeStats()
        rng = np.random.default_rng(0)
        last = None
        for _ in range(20):
            a = rng.integers(0, 28, total)                      # same global action stream on every rank
            env.step(a[lo:hi])
            obs = {k: torch.from_numpy(np.ascontiguousarray(v)) for k, v in env.obs().items()}
            reward, done = torch.from_numpy(env.reward.copy()), torch.from_numpy(env.done.copy())
            stats.update(done, torch.from_numpy(env.flags.copy()), torch.from_numpy(env.fin_return.copy()),
                         torch.from_numpy(env.fin_length.copy()).to(torch.float32))
            last = sharding.all_gather_transitions(obs, reward, done, W)
        g_obs, g_reward, g_done = last
        q.put((rank, g_obs["local_map"].numpy().copy(), g_reward.numpy().copy(), g_done.numpy().copy(),
               stats.all_reduce().numpy().copy()))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_two_rank_gather_and_stats_match_single_process():
    import oracle
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = sorted([q.get(timeout=100) for _ in range(2)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=30)
        assert p.exitcode == 0

    total, W = 8, 16
    env = oracle.OracleVectorEnv(total, grid=32, window=W, seed=5, max_steps=6, block_shift=2)
    env.reset()
    stats = sharding.EpisodeStats()
    rng = np.random.default_rng(0)
    for _ in range(20):
        env.step(rng.integers(0, 28, total))
        stats.update(torch.from_numpy(env.done.copy()), torch.from_numpy(env.flags.copy()),
                     torch.from_numpy(env.fin_return.copy()), torch.from_numpy(env.fin_length.copy()).to(torch.float32))
    for rank, maps, reward, done, red in results:
        assert np.array_equal(maps, env.local_map)              # the gathered global batch == one-process batch
        assert np.array_equal(reward, env.reward) and np.array_equal(done, env.done.astype(bool))
        assert np.allclose(red, stats.acc.numpy())
    assert sharding.EpisodeStats.summary(stats.acc)["episodes"] > 0
