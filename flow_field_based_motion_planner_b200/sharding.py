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


def shard_range(global_envs: int, rank: int, world: int, equal: bool = False):
    """Contiguous shard [lo, hi) of `global_envs` envs for `rank`; the first `global_envs % world` ranks get one more.
    equal=True raises instead: the learner feeds (all_gather_transitions, LearnerFeed) need the same N on every rank."""
    base, extra = divmod(global_envs, world)
    if equal and extra:
        raise ValueError(f"{global_envs} envs do not split evenly over {world} ranks (the learner feeds need equal shards)")
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
    global batch.  One all_gather_into_tensor of world * transition_nbytes bytes.  Every rank must hold the same number of
    envs (shard_range(..., equal=True)): the collective needs equal sizes and the peers' blocks are decoded with the local n."""
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


class _DevicePointer:
    """A library-owned device allocation as a __cuda_array_interface__ object (torch.as_tensor wraps it without a copy)."""

    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 3}


class LearnerFeed:
    """The learner feed as one kernel over NVLink peer memory (csrc/feed.cu, include/ffmp_b200.h) instead of
    pack_transitions + all_gather_into_tensor: every rank's push kernel stores its packed transition block straight into
    the gather buffers of the destination ranks, which map each other's buffers through CUDA IPC.  Only the 64-byte IPC
    handles travel through torch.distributed, once, at construction.

        feed = LearnerFeed(env)                 # on every rank (collective: exchanges the handles)
        feed.push()                             # after env.step(): this rank's block -> every destination
        obs, reward, done = feed.wait()         # zero-copy views [world, N, ...] into the gather buffer (stream ordered)
        ... consume ...
        feed.release()                          # hand the buffer back to the producers

    dest = None pushes to every rank (all-gather semantics); dest = [0] feeds a single learner rank (gather).  A rank that is
    not a destination skips wait() / release()."""

    def __init__(self, env, dest=None, group=None, timeout_s=10.0):
        import ctypes as C
        from . import native
        self._C, self._native, self._L = C, native, native.lib()
        self.env, self.group, self.timeout_s = env, group, float(timeout_s)
        on = dist.is_available() and dist.is_initialized()
        self.world = dist.get_world_size(group) if on else 1
        self.rank = dist.get_rank(group) if on else 0
        self.N, self.W = env.num_envs, env.config.window
        self.block = transition_nbytes(self.N, self.W)
        self.dest = list(range(self.world)) if dest is None else sorted(int(d) for d in dest)
        self.dest_mask = sum(1 << d for d in self.dest)
        self.src_mask = (1 << self.world) - 1
        self._f = C.c_void_p()
        native.check(self._L.ffmp_feed_create(env.device.index, self.world, self.rank, self.block, C.byref(self._f)), "ffmp_feed_create")
        handle = (C.c_uint8 * 64)()
        native.check(self._L.ffmp_feed_handle(self._f, handle), "ffmp_feed_handle")
        handles = [bytes(handle)]
        if self.world > 1:
            both = [None] * self.world
            dist.all_gather_object(both, (bytes(handle), self.N, self.W), group=group)
            if any((n, w) != (self.N, self.W) for _, n, w in both):
                raise ValueError(f"LearnerFeed needs the same num_envs / window on every rank, got {[(n, w) for _, n, w in both]} "
                                 "(shard_range(..., equal=True))")
            handles = [hb for hb, _, _ in both]
        for r, hb in enumerate(handles):
            if r != self.rank:
                buf = (C.c_uint8 * 64).from_buffer_copy(hb)
                native.check(self._L.ffmp_feed_connect(self._f, r, buf), f"ffmp_feed_connect({r})")
        base, slot, bufstride = C.c_void_p(), C.c_size_t(), C.c_size_t()
        native.check(self._L.ffmp_feed_info(self._f, C.byref(base), C.byref(slot), C.byref(bufstride), None), "ffmp_feed_info")
        self.slot_stride, self.buffer_stride = slot.value, bufstride.value
        with torch.cuda.device(env.device):
            self._mem = torch.as_tensor(_DevicePointer(base.value, 2 * bufstride.value), device=env.device)
        self._views = [self._make_views(b) for b in range(2)]
        self.seq = 0
        if self.world > 1:
            dist.barrier(group=group)        # every rank has mapped every buffer before the first push

    def _make_views(self, b):
        n, w, world, ss = self.N, self.W, self.world, self.slot_stride
        ww = w * w
        raw = self._mem
        base = b * self.buffer_stride                      # as_strided offsets are absolute in the storage
        maps = raw.as_strided((world, n, 2, w, w), (ss, 2 * ww, ww, w, 1), base)
        off = base + n * 2 * ww

        def f32(offset, cols):
            v = raw.as_strided((world, n * cols * 4), (ss, 1), offset)
            return v.view(torch.float32).view(world, n, cols) if cols > 1 else v.view(torch.float32)
        rel_goal, velocity, reward = f32(off, 2), f32(off + 8 * n, 2), f32(off + 16 * n, 1)
        done = raw.as_strided((world, n), (ss, 1), off + 20 * n)
        return {"local_map": maps, "relative_goal": rel_goal, "velocity": velocity}, reward, done

    def _stream(self):
        return self._C.c_void_p(torch.cuda.current_stream(self.env.device).cuda_stream)

    def push(self):
        self._native.check(self._L.ffmp_feed_push(self.env._h, self._f, self.dest_mask, self.timeout_s, self._stream()), "ffmp_feed_push")
        self.seq += 1

    def wait(self):
        """(obs, reward, done) of the global batch as views [world, N, ...] (done: uint8); valid until release()."""
        buf = self._C.c_void_p()
        self._native.check(self._L.ffmp_feed_wait(self._f, self.src_mask, self.timeout_s, self._C.byref(buf), self._stream()), "ffmp_feed_wait")
        return self._views[self.seq & 1]

    def release(self):
        self._native.check(self._L.ffmp_feed_release(self._f, self.src_mask, self._stream()), "ffmp_feed_release")

    def error_word(self) -> int:
        w = self._C.c_uint32()
        self._native.check(self._L.ffmp_feed_error(self._f, self._C.byref(w), self._stream()), "ffmp_feed_error")
        return w.value

    def close(self):
        if self._f is not None and self._f.value:
            self._mem = None
            self._views = None
            self._L.ffmp_feed_destroy(self._f)
            self._f = self._C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
