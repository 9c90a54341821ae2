"""Device replay ring — SURVEY.md §8(f) row 4: the reference's `ReplayMemory` / `Transition`
(/root/reference/src/train.py:44, 212-228: a FIFO list of CAPACITY transitions, `push` overwrites the oldest,
`sample` draws a uniform random minibatch) for the batched env, resident on the GPU.

One `push` stores the whole batch's step: the packed transition block the step kernel's outputs form
(`ffmp_pack_transitions`: the two newest ring frames, relative goal, velocity, reward, done of every env, one kernel, 1 B read
+ 1 B written per byte) plus the actions that led to it.  Consecutive blocks share their observations, so a transition
(t, e) is read as state = block[t-1], action / reward / done / next state = block[t] — the ten `Transition` fields without
storing any observation twice.  `state_t` / `observe_t` (the odometry dt, train.py:553-557) is the env's constant `dt`.
After an episode end the stored next observation is the first one of the next episode (auto-reset, SPEC.md §7); `done` marks
those transitions, as in any vectorised gym env.
"""
import ctypes as C

import torch

from . import native
from .sharding import transition_nbytes


class ReplayRing:
    def __init__(self, env, capacity_steps: int):
        if capacity_steps < 2:
            raise ValueError("capacity_steps must be >= 2 (a transition spans two consecutive steps)")
        self.env, self.T = env, int(capacity_steps)
        self.N, self.W = env.num_envs, env.config.window
        self.block = transition_nbytes(self.N, self.W)
        self.stride = (self.block + 255) // 256 * 256
        self._L = native.lib()
        dev = env.device
        self.blocks = torch.zeros((self.T, self.stride), dtype=torch.uint8, device=dev)
        self.actions = torch.zeros((self.T, self.N), dtype=torch.int64, device=dev)
        self.count = 0                      # pushes so far; slot of push k is k % T
        n, ww = self.N, self.W * self.W
        raw = self.blocks
        base = raw.storage_offset()
        self.maps = raw.as_strided((self.T, n, 2, self.W, self.W), (self.stride, 2 * ww, ww, self.W, 1), base)
        off = base + n * 2 * ww

        def f32(offset, cols):
            v = raw.as_strided((self.T, n * cols * 4), (self.stride, 1), offset).view(torch.float32)
            return v.view(self.T, n, cols) if cols > 1 else v
        self.rel_goal, self.velocity, self.reward = f32(off, 2), f32(off + 8 * n, 2), f32(off + 16 * n, 1)
        self.done = raw.as_strided((self.T, n), (self.stride, 1), off + 20 * n)

    def __len__(self):
        """Number of stored transitions (per env times the number of envs)."""
        return max(0, min(self.count, self.T) - 1) * self.N

    def push(self, actions=None):
        """Store the env's current step (call after reset() with actions=None, then after every step(actions))."""
        slot = self.count % self.T
        stream = C.c_void_p(torch.cuda.current_stream(self.env.device).cuda_stream)
        with torch.cuda.device(self.env.device):
            native.check(self._L.ffmp_pack_transitions(self.env._h, C.c_void_p(self.blocks[slot].data_ptr()), stream),
                         "ffmp_pack_transitions")
            if actions is not None:
                self.actions[slot].copy_(actions.reshape(self.N))
        self.count += 1

    def sample(self, batch_size: int, generator=None):
        """Uniform random minibatch of stored transitions (ReplayMemory.sample, train.py:224-225) as device tensors:
        state_m u8[B,2,W,W], state_g, state_v f32[B,2], action i64[B], observe_m, observe_g, observe_v, reward f32[B], done u8[B]."""
        stored = min(self.count, self.T)
        if stored < 2:
            raise ValueError("the ring holds no complete transition yet")
        dev = self.env.device
        newest = self.count - 1
        # transition k (newest - stored + 2 <= k <= newest) pairs push k-1 with push k
        k = newest - torch.randint(0, stored - 1, (batch_size,), device=dev, generator=generator)
        e = torch.randint(0, self.N, (batch_size,), device=dev, generator=generator)
        t1, t0 = k % self.T, (k - 1) % self.T
        return {"state_m": self.maps[t0, e], "state_g": self.rel_goal[t0, e], "state_v": self.velocity[t0, e],
                "action": self.actions[t1, e], "observe_m": self.maps[t1, e], "observe_g": self.rel_goal[t1, e],
                "observe_v": self.velocity[t1, e], "reward": self.reward[t1, e], "done": self.done[t1, e],
                "index": torch.stack([k, e], 1)}

    def sample_learner(self, batch_size: int, generator=None, out=None):
        """The same uniform minibatch in the learner's formats, gathered by ONE kernel (csrc/replay.cu): observation stacks as
        bf16 NCHW [B,2,W,W] (the Q network's input; the u8 -> bf16 widening is fused into the gather), state_g / state_v /
        observe_g / observe_v f32 [B,2], reward f32 [B], done u8 [B], action i64 [B].  `out` reuses a previous result's tensors."""
        stored = min(self.count, self.T)
        if stored < 2:
            raise ValueError("the ring holds no complete transition yet")
        dev = self.env.device
        newest = self.count - 1
        k = newest - torch.randint(0, stored - 1, (batch_size,), device=dev, generator=generator)
        e = torch.randint(0, self.N, (batch_size,), device=dev, generator=generator)
        index = torch.stack([k, e], 1).contiguous()
        B, W = batch_size, self.W
        if out is None:
            out = {"state_m": torch.empty((B, 2, W, W), dtype=torch.bfloat16, device=dev),
                   "observe_m": torch.empty((B, 2, W, W), dtype=torch.bfloat16, device=dev),
                   "state_g": torch.empty((B, 2), device=dev), "state_v": torch.empty((B, 2), device=dev),
                   "observe_g": torch.empty((B, 2), device=dev), "observe_v": torch.empty((B, 2), device=dev),
                   "reward": torch.empty((B,), device=dev), "done": torch.empty((B,), dtype=torch.uint8, device=dev),
                   "action": torch.empty((B,), dtype=torch.int64, device=dev)}
        stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        p = lambda t: C.c_void_p(t.data_ptr())      # noqa: E731
        rc = self._L.ffmp_replay_gather(dev.index, p(self.blocks), self.stride, self.T, self.N, W, p(self.actions), p(index), B,
                                        p(out["state_m"]), p(out["observe_m"]), p(out["state_g"]), p(out["state_v"]),
                                        p(out["observe_g"]), p(out["observe_v"]), p(out["reward"]), p(out["done"]),
                                        p(out["action"]), stream)
        native.check(rc, "ffmp_replay_gather")
        out["index"] = index
        return out
