"""Multi-GPU plumbing: env sharding by global id, learner feed (all-gather) and episode statistics
(all-reduce).  One process per GPU (torchrun); NCCL over NVLink on GPUs, gloo in the CPU tests.

Environments are independent: rank r owns global env ids [r*N, (r+1)*N) and seeds derive
from the global id, so results do not depend on the GPU count and the simulation itself needs NO
collective.  Only two exchanges exist:
  * learner feed: the per-step (or per-rollout) transition block [local_map u8 | relative_goal, velocity
    f32 | reward f32 | done u8] of every rank, all-gathered as one byte tensor;
  * episode statistics: one all-reduce(SUM) of five numbers at log cadence.
This module only moves tensors (torch.distributed); it contains no simulation arithmetic.
"""
from dataclasses import dataclass

import torch
import torch.distributed as dist


def shard_range(global_envs: int, rank: int, world: int):
    """Contiguous shard [lo, hi) of `global_envs` envs for `rank`; the first `global_envs % world` ranks get one more."""
    base, extra = divmod(global_envs, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def transition_nbytes(num_envs: int, window: int) -> int:
    return num_envs * (2 * window * window + 8 + 8 + 4 + 1)


def pack_transitions(obs, reward, done, out=None):
    """Pack one step of one shard into a flat uint8 tensor (layout: maps | rel_goal | velocity | reward | done)."""
    lm = obs["local_map"]
    n = lm.shape[0]
    parts = [lm.reshape(n, -1), obs["relative_goal"].contiguous().view(torch.uint8).reshape(n, -1),
             obs["velocity"].contiguous().view(torch.uint8).reshape(n, -1),
             reward.contiguous().view(torch.uint8).reshape(n, -1), done.to(torch.uint8).reshape(n, 1)]
    sizes = [p.shape[1] for p in parts]
    total = n * sum(sizes)
    if out is None:
        out = torch.empty(total, dtype=torch.uint8, device=lm.device)
    off = 0
    for p, s in zip(parts, sizes):
        out[off:off + n * s].view(n, s).copy_(p)
        off += n * s
    return out


def unpack_transitions(buf, num_envs: int, window: int):
    n, ww = num_envs, window * window
    off = 0

    def take(nbytes):
        nonlocal off
        v = buf[off:off + n * nbytes].view(n, nbytes)
        off += n * nbytes
        return v

    maps = take(2 * ww).view(n, 2, window, window)
    rel_goal = take(8).contiguous().view(torch.float32).view(n, 2)
    velocity = take(8).contiguous().view(torch.float32).view(n, 2)
    reward = take(4).contiguous().view(torch.float32).view(n)
    done = take(1).view(n) != 0
    return {"local_map": maps, "relative_goal": rel_goal, "velocity": velocity}, reward, done


def all_gather_transitions(obs, reward, done, window: int, group=None):
    """Every rank receives every shard's transition block, in rank order.  Returns (obs, reward, done) of the
    global batch.  One all_gather_into_tensor of world * transition_nbytes bytes."""
    n = obs["local_map"].shape[0]
    local = pack_transitions(obs, reward, done)
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return unpack_transitions(local, n, window)
    world = dist.get_world_size(group)
    gathered = torch.empty(world * local.numel(), dtype=torch.uint8, device=local.device)
    dist.all_gather_into_tensor(gathered, local, group=group)
    per = local.numel()
    outs = [unpack_transitions(gathered[r * per:(r + 1) * per], n, window) for r in range(world)]
    obs_g = {k: torch.cat([o[0][k] for o in outs]) for k in outs[0][0]}
    return obs_g, torch.cat([o[1] for o in outs]), torch.cat([o[2] for o in outs])


@dataclass
class EpisodeStats:
    """Running sums of finished episodes: [sum_return, sum_length, n_episodes, n_goal, n_collision]."""
    acc: torch.Tensor = None

    def update(self, done, flags, fin_return, fin_length):
        d = done.to(torch.bool)
        row = torch.stack([
            (fin_return * d).sum(dtype=torch.float64),
            (fin_length * d).sum(dtype=torch.float64),
            d.sum(dtype=torch.float64),
            (d & ((flags & 2) != 0)).sum(dtype=torch.float64),
            (d & ((flags & 1) != 0)).sum(dtype=torch.float64),
        ])
        self.acc = row if self.acc is None else self.acc + row
        return self

    def all_reduce(self, group=None):
        """Global sums over all ranks (one 5-element all-reduce)."""
        t = self.acc.clone()
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
        return t

    @staticmethod
    def summary(t):
        n = max(float(t[2]), 1.0)
        return {"episodes": int(t[2]), "mean_return": float(t[0]) / n, "mean_length": float(t[1]) / n,
                "reach_rate": float(t[3]) / n, "collision_rate": float(t[4]) / n}
