"""FFMPVectorEnv — batched, GPU-resident `FFMP-v0` with the reference's gym surface.

Host side only: allocation of device buffers (torch), the ctypes calls into libffmp_b200.so and
tensor views.  All arithmetic happens in the CUDA kernels (csrc/); there is no CPU fallback.

Reference surface mirrored (paths under /root/reference/src/):
  * spaces               gym_ffmp/envs/ffmp.py:28-64
  * reward / done rules  gym_ffmp/envs/ffmp.py:85-188   (call site train.py:577)
  * observation layout   train.py:474-486, 535-557      (maps NCHW oldest-first, rel goal, velocity)
  * episode sequencing   train.py:559-565, 593, 607-608, 611-682
  * gym (Dec 2020) API   reset() -> obs ; step(a) -> (obs, reward, done, info)
"""
import ctypes as C
import math
from dataclasses import dataclass

import numpy as np
import torch

from . import native, spaces
from .robot import NUM_ACTIONS, RobotAction

MAP_RANGE = 5.0          # ffmp.py:14
MAP_GRID_NUM = 100       # ffmp.py:15
MAP_RESOLUTION = 0.05    # ffmp.py:18
ROBOT_RSIZE = 0.13       # ffmp.py:17
GOAL_THRESHOLD = 0.5     # ffmp.py:19
MAX_STEPS = 200          # train.py:60


def p_threshold(p_occ: float) -> int:
    """SPEC.md §3: min(2^32-1, floor(p_occ * 2^32)) in fp64."""
    return int(min(4294967295, math.floor(float(p_occ) * 4294967296.0)))


@dataclass
class FFMPConfig:
    num_envs: int = 1
    grid: int = 128                 # G
    window: int = MAP_GRID_NUM      # W (local map side)
    ring: int = 8                   # K observation frame slots (2 = contiguous [N,2,W,W])
    slots: int = 8                  # S resident scenario slots per env (2..16)
    regen_batch: int = 0            # ticks per background regeneration launch (0 = library default, see ffmp_b200.h)
    max_steps: int = MAX_STEPS
    goal_mode: int = 0              # 0 re-sampled per episode, 1 static at (G-8, G-8)
    block_shift: int = 3
    p_occ: float = 0.10
    seed: int = 0
    env_id_base: int = 0
    dt: float = 0.1
    terminal_obs: bool = False      # also keep the terminal local_map of finished envs (info["terminal_local_map"])
    device: str = "cuda:0"


def make_spaces(window: int = MAP_GRID_NUM):
    """Per-env spaces, equal to FFMP.__init__ (ffmp.py:28-64)."""
    act = RobotAction()
    action_low = np.array([act.cmd[0].linear_v, act.cmd[0].angular_v])
    action_high = np.array([act.cmd[27].linear_v, act.cmd[27].angular_v])
    action_space = spaces.Box(action_low, action_high, dtype=np.float32)
    obs = {
        "local_map": spaces.Box(np.full((window, window, 1), 0), np.full((window, window, 1), 255), dtype=np.int32),
        "relative_goal": spaces.Box(np.array([0.0, 0.0]), np.array([math.sqrt(2.0) * MAP_RANGE, math.pi]), dtype=np.float32),
        "velocity": spaces.Box(action_low, action_high, dtype=np.float32),
    }
    return action_space, spaces.Dict(obs), spaces.Dict(dict(obs))


class FFMPVectorEnv:
    """N independent FFMP environments stepped by one fused CUDA kernel per call."""

    def __init__(self, num_envs=None, config: FFMPConfig = None, **kwargs):
        cfg = config or FFMPConfig()
        if num_envs is not None:
            kwargs["num_envs"] = num_envs
        if kwargs:
            cfg = FFMPConfig(**{**cfg.__dict__, **kwargs})
        self.config = cfg
        self._L = native.lib()                      # raises if the CUDA library is missing
        if not torch.cuda.is_available():
            raise native.NativeError("FFMPVectorEnv needs a CUDA device (B200); there is no CPU fallback")
        self.device = torch.device(cfg.device)
        if self.device.type != "cuda":
            raise native.NativeError("FFMPVectorEnv device must be a CUDA device")
        dev_index = self.device.index if self.device.index is not None else torch.cuda.current_device()
        self.device = torch.device("cuda", dev_index)
        self.num_envs = cfg.num_envs
        self.action_space, self.observation_space, self.state_space = make_spaces(cfg.window)
        self.single_action_space, self.single_observation_space = self.action_space, self.observation_space
        self.discrete_action_space = spaces.Discrete(NUM_ACTIONS)   # what the agent actually emits (train.py:70,345)
        self.action = RobotAction()

        c = native.Cfg(abi_version=native.ABI_VERSION, device=dev_index, num_envs=cfg.num_envs, grid=cfg.grid,
                       window=cfg.window, ring=cfg.ring, slots=cfg.slots, max_steps=cfg.max_steps,
                       goal_mode=cfg.goal_mode, block_shift=cfg.block_shift, p_thresh=p_threshold(cfg.p_occ),
                       env_id_base=cfg.env_id_base, seed=cfg.seed, dt=cfg.dt, regen_batch=cfg.regen_batch)
        self._cfg = c
        sz = native.Sizes()
        native.check(self._L.ffmp_query_sizes(C.byref(c), C.byref(sz)), "ffmp_query_sizes")
        self._h = C.c_void_p()
        native.check(self._L.ffmp_create(C.byref(c), C.byref(self._h)), "ffmp_create")

        N, G, W, K, S = cfg.num_envs, cfg.grid, cfg.window, cfg.ring, cfg.slots
        d = self.device
        with torch.cuda.device(d):
            self.cost = torch.zeros((S, N, G, G), dtype=torch.int32, device=d)
            self.flow = torch.zeros((S, N, G, G), dtype=torch.uint8, device=d)
            self.scen = torch.zeros((S, N, 8), dtype=torch.int32, device=d)
            self.state = torch.zeros((N, 16), dtype=torch.int32, device=d)
            self.frames = torch.zeros((N, K, W, W), dtype=torch.uint8, device=d)
            # the five small per-step outputs share one block (reward | rel_goal | velocity | done | flags) so that
            # ffmp_step_host can bring them to the host with a single copy (include/ffmp_b200.h)
            # (N even keeps the float2 stores of the kernels 8-byte aligned; odd N falls back to separate buffers)
            if N % 2 == 0:
                self._out_block = torch.zeros((22 * N + 16,), dtype=torch.uint8, device=d)
                self.reward, self.rel_goal, self.velocity, self.done, self.flags = self._split_out_block(self._out_block, N)
            else:
                self.reward = torch.zeros((N,), dtype=torch.float32, device=d)
                self.rel_goal = torch.zeros((N, 2), dtype=torch.float32, device=d)
                self.velocity = torch.zeros((N, 2), dtype=torch.float32, device=d)
                self.done = torch.zeros((N,), dtype=torch.uint8, device=d)
                self.flags = torch.zeros((N,), dtype=torch.uint8, device=d)
            self.term_rel_goal = torch.zeros((N, 2), dtype=torch.float32, device=d)
            self.term_velocity = torch.zeros((N, 2), dtype=torch.float32, device=d)
            self.fin_return = torch.zeros((N,), dtype=torch.float32, device=d)
            self.fin_length = torch.zeros((N,), dtype=torch.int32, device=d)
            self._workspace = torch.zeros((sz.workspace,), dtype=torch.uint8, device=d)
        assert self.flow.numel() == sz.flow and self.cost.numel() * 4 == sz.cost and self.frames.numel() == sz.frames
        b = native.Buffers(**{name: getattr(self, name).data_ptr() for name in
                              ("cost", "flow", "scen", "state", "frames", "rel_goal", "velocity", "reward",
                               "done", "flags", "term_rel_goal", "term_velocity", "fin_return", "fin_length")},
                           workspace=self._workspace.data_ptr())
        native.check(self._L.ffmp_bind(self._h, C.byref(b)), "ffmp_bind")
        self.term_local_map = None
        if cfg.terminal_obs:
            # rows of envs that did not finish keep their last terminal observation (train.py:611-664 reads it on `done` only)
            with torch.cuda.device(d):
                self.term_local_map = torch.zeros((N, 2, W, W), dtype=torch.uint8, device=d)
            native.check(self._L.ffmp_set_terminal_obs(self._h, C.c_void_p(self.term_local_map.data_ptr())), "ffmp_set_terminal_obs")
        self._done_bool = self.done.view(torch.bool)
        self._views = {}
        self._host_out = {}
        self._host = None
        self._is_reset = False

    # ------------------------------------------------------------------------------------------
    @staticmethod
    def _split_out_block(block, N):
        reward = block[0:4 * N].view(torch.float32)
        rel_goal = block[4 * N:12 * N].view(torch.float32).view(N, 2)
        velocity = block[12 * N:20 * N].view(torch.float32).view(N, 2)
        done = block[20 * N:21 * N]
        flags = block[21 * N:22 * N]
        return reward, rel_goal, velocity, done, flags

    def _stream(self):
        # the raw handle of torch's current stream on the env's device (torch.cuda.current_stream() builds a Stream object: ~1 us)
        return C.c_void_p(torch._C._cuda_getCurrentRawStream(self.device.index))

    def _obs(self):
        p = C.c_int32()
        native.check(self._L.ffmp_obs_slot(self._h, C.byref(p)), "ffmp_obs_slot")
        view = self._views.get(p.value)
        if view is None:
            view = self.frames.narrow(1, p.value - 1, 2)      # [N,2,W,W] = [older, newest] (train.py:474-486)
            self._views[p.value] = view
        return {"local_map": view, "relative_goal": self.rel_goal, "velocity": self.velocity}

    def seed(self, seed=None):
        """Re-seed; takes effect at the next full reset()."""
        if seed is not None:
            self.close()
            self.__init__(config=FFMPConfig(**{**self.config.__dict__, "seed": int(seed)}))
        return [self.config.seed]

    def reset(self, mask=None):
        """reset() restarts every env from episode 0 of its seed; reset(mask) starts the next episode of
        the masked envs (bool/uint8 tensor on the device)."""
        with torch.cuda.device(self.device):
            if mask is None:
                native.check(self._L.ffmp_reset(self._h, None, self._stream()), "ffmp_reset")
                self._is_reset = True
            else:
                m = torch.as_tensor(mask, device=self.device).to(torch.uint8).contiguous()
                native.check(self._L.ffmp_reset(self._h, C.c_void_p(m.data_ptr()), self._stream()), "ffmp_reset(mask)")
        return self._obs()

    def step(self, actions):
        """actions: int64[N] action ids on the device -> (obs, reward f32[N], done bool[N], info)."""
        a = actions
        if not (isinstance(a, torch.Tensor) and a.device == self.device and a.dtype == torch.int64 and a.is_contiguous()):
            a = torch.as_tensor(a).to(device=self.device, dtype=torch.int64).contiguous()
        if a.numel() != self.num_envs:
            raise ValueError(f"expected {self.num_envs} actions, got {a.numel()}")
        with torch.cuda.device(self.device):
            native.check(self._L.ffmp_step(self._h, C.c_void_p(a.data_ptr()), self._stream()), "ffmp_step")
        return self._obs(), self.reward, self._done_bool, self._info()

    def _info(self):
        info = {"flags": self.flags, "terminal_relative_goal": self.term_rel_goal,
                "terminal_velocity": self.term_velocity, "episode_return": self.fin_return,
                "episode_length": self.fin_length, "observe_t": self.config.dt}      # observe_t: train.py:553-557 (constant dt)
        if self.term_local_map is not None:
            info["terminal_local_map"] = self.term_local_map
        return info

    @staticmethod
    def decode_flags(flags):
        return {"is_collision": (flags & 1) != 0, "is_goal": (flags & 2) != 0, "truncated": (flags & 4) != 0}

    def rollout(self, actions, graph=False):
        """actions int64[T,N] on the device: T back-to-back steps with no host work in between.  graph=True replays the T
        steps as one CUDA graph launch (captured the first time this action buffer, T and ring phase are seen; the buffer's
        contents are re-read by every replay, so refill it in place)."""
        if not actions.is_contiguous():
            if graph:
                raise ValueError("rollout(graph=True) needs a contiguous action tensor (its address is part of the graph)")
            actions = actions.contiguous()
        a = actions
        assert a.dtype == torch.int64 and a.device == self.device and a.shape[1] == self.num_envs
        fn = self._L.ffmp_rollout_graphed if graph else self._L.ffmp_rollout
        with torch.cuda.device(self.device):
            native.check(fn(self._h, C.c_void_p(a.data_ptr()), int(a.shape[0]), self._stream()), "ffmp_rollout")
        if graph:
            self._graph_actions = a          # keeps the captured buffer alive
        return self._obs(), self.reward, self._done_bool, self._info()

    def _host_buffers(self):
        N = self.num_envs
        if self._host is None:
            if N % 2 == 0:
                block = torch.zeros((22 * N + 16,), dtype=torch.uint8, pin_memory=True)
                r, g, v, dn, fl = self._split_out_block(block, N)
            else:
                pin = dict(pin_memory=True)
                r, g, v = torch.empty(N, dtype=torch.float32, **pin), torch.empty((N, 2), dtype=torch.float32, **pin), torch.empty((N, 2), dtype=torch.float32, **pin)
                dn, fl = torch.empty(N, dtype=torch.uint8, **pin), torch.empty(N, dtype=torch.uint8, **pin)
            self._host = {"reward": r, "rel_goal": g, "velocity": v, "done": dn, "flags": fl, "done_bool": dn.view(torch.bool),
                          "info": {"flags": fl}}
            self._host_ptrs = tuple(C.c_void_p(self._host[k].data_ptr()) for k in ("reward", "done", "flags", "rel_goal", "velocity"))
        return self._host

    def step_async(self, actions_host):
        """gym.vector's step_async with host buffers: queue one step on actions int64[N] in (pinned) host memory and
        return at once; the buffer must stay untouched until step_wait() returns."""
        if self._host is None:
            self._host_buffers()
        a = actions_host
        if a.dtype is not torch.int64 or a.numel() != self.num_envs or a.device.type != "cpu" or not a.is_contiguous():
            raise ValueError("step_async expects a contiguous int64 CPU tensor with one action id per env")
        # the library switches to its own device (DeviceGuard): no torch device context here
        rc = self._L.ffmp_step_host_async(self._h, a.data_ptr(), *self._host_ptrs,
                                          torch._C._cuda_getCurrentRawStream(self.device.index))
        if rc:
            native.check(rc, "ffmp_step_host_async")
        self._pending_actions = a        # keeps the buffer alive while the kernel reads it

    def step_wait(self):
        """gym.vector's step_wait: block until the step queued by step_async has delivered -> (obs, reward, done, info)
        with reward, done, flags, relative_goal and velocity as pinned host tensors; local_map stays on the device.
        The returned containers are cached per frame slot (do not mutate them)."""
        p = C.c_int32()
        self._L.ffmp_obs_slot(self._h, C.byref(p))                  # host bookkeeping done while the GPU works
        out = self._host_out.get(p.value)
        if out is None:
            hst = self._host
            obs = {"local_map": self._obs()["local_map"], "relative_goal": hst["rel_goal"], "velocity": hst["velocity"]}
            out = self._host_out[p.value] = (obs, hst["reward"], hst["done_bool"], hst["info"])
        rc = self._L.ffmp_step_host_wait(self._h)
        if rc:
            native.check(rc, "ffmp_step_host_wait")
        self._pending_actions = None
        return out

    def step_host(self, actions_host):
        """Host-buffer step: actions int64[N] in (pinned) host memory -> (obs, reward, done, info) with reward,
        done, flags, relative_goal and velocity returned as host tensors; local_map stays on the device.
        With pinned buffers the results arrive without a device-to-host copy or a stream synchronisation (include/ffmp_b200.h)."""
        self.step_async(actions_host)
        return self.step_wait()

    @property
    def h2d_bytes_per_step(self):
        return self.num_envs * 8

    @property
    def d2h_bytes_per_step(self):
        return self.num_envs * (4 + 1 + 1 + 8 + 8)

    def learner_input(self, dtype=torch.float32, scale=1.0, out=None):
        """The current observation as the learner's input: float32 / bfloat16 [N, 2, W, W], oldest frame first (the
        reference's `observe_m`, train.py:474-486, 539-545; scale=1.0 is its plain `.float()`)."""
        if dtype not in (torch.float32, torch.bfloat16):
            raise ValueError("dtype must be torch.float32 or torch.bfloat16")
        N, W = self.num_envs, self.config.window
        if out is None:
            out = torch.empty((N, 2, W, W), dtype=dtype, device=self.device)
        assert out.dtype == dtype and out.device == self.device and out.is_contiguous() and out.numel() == N * 2 * W * W
        with torch.cuda.device(self.device):
            native.check(self._L.ffmp_learner_input(self._h, C.c_void_p(out.data_ptr()), 0 if dtype == torch.float32 else 1,
                                                    C.c_float(scale), self._stream()), "ffmp_learner_input")
        return out

    def scan(self, beams=360, range_max=3.5, out=None):
        """LiDAR scan of every env at its current pose on its current scenario (SPEC.md §9; the `/scan` input of
        FFMP.rewarder2, train.py:87,144-150,577) -> (ranges f32[N,beams], hit u8[N] = is_collision2 of the beams)."""
        N = self.num_envs
        if out is None:
            out = torch.empty((N, beams), dtype=torch.float32, device=self.device)
        assert out.dtype == torch.float32 and out.device == self.device and out.is_contiguous() and out.numel() == N * beams
        hit = torch.empty((N,), dtype=torch.uint8, device=self.device)
        with torch.cuda.device(self.device):
            native.check(self._L.ffmp_scan(self._h, int(beams), C.c_float(range_max), C.c_void_p(out.data_ptr()),
                                           C.c_void_p(hit.data_ptr()), self._stream()), "ffmp_scan")
        return out, hit

    def join(self):
        """Order the current stream after all queued background scenario regeneration."""
        with torch.cuda.device(self.device):
            native.check(self._L.ffmp_join(self._h, self._stream()), "ffmp_join")

    def kernel_timing(self, enable: bool):
        """enable=True: start per-kernel CUDA-event timing of the next <=256 ticks; enable=False: stop and return
        {'tick_ms', 'regen_ms', 'ticks'} (averages per launch: the step kernel, the background regeneration launch)."""
        d, o, n = C.c_float(), C.c_float(), C.c_int32()
        with torch.cuda.device(self.device):
            native.check(self._L.ffmp_timing(self._h, 1 if enable else 0, C.byref(d), C.byref(o), C.byref(n)), "ffmp_timing")
        return None if enable else {"tick_ms": d.value, "regen_ms": o.value, "ticks": n.value}

    def launch_count(self) -> int:
        """Kernels launched by this env since construction (step, reset and background regeneration launches)."""
        n = C.c_uint64()
        native.check(self._L.ffmp_launch_count(self._h, C.byref(n)), "ffmp_launch_count")
        return n.value

    def error_word(self) -> int:
        w = C.c_uint32()
        with torch.cuda.device(self.device):
            native.check(self._L.ffmp_error_word(self._h, C.byref(w), self._stream()), "ffmp_error_word")
        return w.value

    # ---- inspection helpers (parity tests, debugging) ------------------------------------------
    def pose(self):
        return self.state[:, 0:3].view(torch.float32)

    def goal(self):
        return self.state[:, 3:5].view(torch.float32)

    def episode(self):
        return self.state[:, 8]

    def steps(self):
        return self.state[:, 7]

    def _current(self, planes):
        slot = (self.episode().to(torch.int64) % self.config.slots)
        idx = torch.arange(self.num_envs, device=self.device)
        return planes[slot, idx]

    def cost_field(self):
        """Integration field of every env's current scenario, int32 [N,G,G]."""
        self.join()
        return self._current(self.cost)

    def flow_image(self):
        self.join()
        return self._current(self.flow)

    def occupancy(self):
        """Occupancy of every env's current scenario (uint8 0/1), decoded from the flow image (255 = occupied)."""
        return (self.flow_image() == 255).to(torch.uint8)

    def flow_dir(self):
        """Direction codes 0..7, 8 = none (SPEC.md §5 decoding of the flow image)."""
        img = self.flow_image()
        return torch.where(img == 255, torch.full_like(img, 8), img // 28)

    # ---- checkpoint / resume of the env (env checkpoint = the state tensors + RNG counters) -----------
    _STATE_TENSORS = ("state", "scen", "flow", "cost", "frames", "reward", "rel_goal", "velocity", "done", "flags",
                      "term_rel_goal", "term_velocity", "fin_return", "fin_length")

    def state_dict(self, planes=True):
        """Everything a later load_state_dict needs to continue bit-identically: the per-env state records, the resident
        scenario slots (every random draw is keyed by (seed, global env id, episode), so there is no RNG state beyond the
        episode counters), the frame ring and the library's step counter.  planes=False leaves out the flow / cost planes
        (S*N*G*G*5 bytes); load_state_dict then regenerates them from the scenario keys."""
        self.join()
        torch.cuda.synchronize(self.device)
        p = C.c_int32()
        native.check(self._L.ffmp_obs_slot(self._h, C.byref(p)), "ffmp_obs_slot")
        names = [n for n in self._STATE_TENSORS if planes or n not in ("flow", "cost")]
        return {"config": dict(self.config.__dict__), "newest_slot": p.value, "planes": bool(planes),
                "tensors": {n: getattr(self, n).detach().cpu().clone() for n in names}}

    def load_state_dict(self, sd):
        """Restore a state_dict taken from an env of the same configuration (device may differ)."""
        want = {k: v for k, v in sd["config"].items() if k != "device"}
        have = {k: v for k, v in self.config.__dict__.items() if k != "device"}
        if want != have:
            raise ValueError(f"checkpoint config {want} does not match this env {have}")
        self.reset()                                   # arms the library (lists, counters) and regenerates every slot
        self.join()
        for n, t in sd["tensors"].items():
            getattr(self, n).copy_(t.to(self.device))
        if not sd.get("planes", True):
            self._regenerate_planes()
        native.check(self._L.ffmp_set_obs_slot(self._h, int(sd["newest_slot"])), "ffmp_set_obs_slot")
        torch.cuda.synchronize(self.device)
        return self

    def _regenerate_planes(self):
        """flow / cost planes of every resident slot from the scenario records' episode keys (checkpoint without planes)."""
        from . import ops
        S, N, G = self.config.slots, self.num_envs, self.config.grid
        ep = self.episode().to(torch.int64)                                       # current episode k of every env
        gids = torch.arange(N, device=self.device) + self.config.env_id_base
        for s in range(S):
            # slot s holds the episode e >= k with e % S == s (the library regenerates a consumed slot for episode k + S - 1)
            e = ep + ((s - ep) % S)
            occ, scen = ops.generate_scenarios(gids, e, G, p_occ=self.config.p_occ, goal_mode=self.config.goal_mode,
                                               block_shift=self.config.block_shift, seed=self.config.seed)
            cost, flow = ops.flow_field(occ, scen[:, 5:7].contiguous())
            self.flow[s].copy_(flow)
            self.cost[s].copy_(cost)

    def close(self):
        h = getattr(self, "_h", None)
        if h is not None and h.value:
            self._L.ffmp_destroy(h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
