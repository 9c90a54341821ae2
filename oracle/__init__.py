"""CPU oracle for the gym_ffmp hot path — TEST INFRASTRUCTURE ONLY.

ctypes front-end over oracle/ffmp_oracle.c (see that file's header).  Importers allowed:
tests/, __graft_entry__.smoke(), bench.py's cpu_baseline / --impl reference legs.  The product
package never imports this module.

Parity status: reference-pinned (tests/golden/ref_golden.json) for the action table, spaces,
footprint/collision, goal test, reward, done, pi_to_pi, relative goal, velocity and the
2-frame stack; PARITY UNPINNED for the [SPEC] parts the reference does not implement
(map generator, integration field, flow direction, kinematics, crop, step/reset sequencing).
"""
import ctypes as C
import math

import numpy as np

from .build import build

_lib = None


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(build())
        f64p, f32p = C.POINTER(C.c_double), C.POINTER(C.c_float)
        u8p, i32p, u32p, i64p = C.POINTER(C.c_uint8), C.POINTER(C.c_int32), C.POINTER(C.c_uint32), C.POINTER(C.c_int64)
        sig = {
            "orc_ref_action": (None, [C.c_int, f64p, f64p]),
            "orc_ref_is_collision": (C.c_int, [i32p, C.c_int, C.c_double, C.c_double, C.c_double]),
            "orc_ref_footprint": (C.c_int, [C.c_int, C.c_double, C.c_double, C.c_double, i32p]),
            "orc_ref_is_collision2": (C.c_int, [f64p, C.c_int]),
            "orc_ref_is_goal": (C.c_int, [C.c_double]),
            "orc_ref_reward": (C.c_double, [C.c_double, C.c_int, C.c_int, C.c_int, f64p]),
            "orc_ref_is_done": (C.c_int, [C.c_int, C.c_int]),
            "orc_ref_pi_to_pi": (C.c_double, [C.c_double]),
            "orc_ref_relative_goal": (None, [C.c_double] * 5 + [f64p]),
            "orc_ref_velocity": (None, [C.c_double] * 6 + [f64p]),
            "orc_pi_to_pi": (C.c_float, [C.c_float]),
            "orc_sincos": (None, [C.c_float, f32p, f32p]),
            "orc_atan2": (C.c_float, [C.c_float, C.c_float]),
            "orc_dist": (C.c_float, [C.c_float, C.c_float]),
            "orc_mix32": (C.c_uint32, [C.c_uint32]),
            "orc_key": (C.c_uint32, [C.c_uint64, C.c_uint32, C.c_uint32]),
            "orc_draw": (C.c_uint32, [C.c_uint32, C.c_uint32, C.c_uint32]),
            "orc_scenario": (None, [C.c_uint64, C.c_uint32, C.c_uint32, C.c_int, C.c_uint32, C.c_int, C.c_int, u8p, f32p, f32p, i32p]),
            "orc_integration_field": (None, [u8p, C.c_int, C.c_int, C.c_int, i32p]),
            "orc_flow_dir": (None, [u8p, i32p, C.c_int, u8p]),
            "orc_flow_image": (None, [u8p, u8p, C.c_int, u8p]),
            "orc_flow_field": (None, [u8p, C.c_int, C.c_int, C.c_int, i32p, u8p, u8p]),
            "orc_crop": (None, [u8p, C.c_int, C.c_int, C.c_int, C.c_int, u8p]),
            "orc_collision": (C.c_int, [u8p, C.c_int, C.c_int, C.c_int]),
            "orc_scan": (C.c_int, [u8p, C.c_int, C.c_int, C.c_float, C.c_float, C.c_float, C.c_int, C.c_float, f32p]),
            "orc_env_create": (C.c_void_p, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_uint32, C.c_uint64, C.c_uint32, C.c_float, C.c_int]),
            "orc_env_destroy": (None, [C.c_void_p]),
            "orc_env_reset": (None, [C.c_void_p]),
            "orc_env_reset_masked": (None, [C.c_void_p, u8p]),
            "orc_env_step": (None, [C.c_void_p, i64p]),
            "orc_env_error_word": (C.c_uint32, [C.c_void_p]),
        }
        for name, (res, args) in sig.items():
            fn = getattr(L, name)
            fn.restype, fn.argtypes = res, args
        for name, ty in _ENV_FIELDS.items():
            fn = getattr(L, "orc_env_" + name)
            fn.restype, fn.argtypes = C.POINTER(ty[0]), [C.c_void_p]
        _lib = L
    return _lib


_ENV_FIELDS = {
    "occ": (C.c_uint8, np.uint8), "dir": (C.c_uint8, np.uint8), "flow": (C.c_uint8, np.uint8),
    "cost": (C.c_int32, np.int32), "pose": (C.c_float, np.float32), "goal": (C.c_float, np.float32),
    "d_first": (C.c_float, np.float32), "ep_return": (C.c_float, np.float32), "steps": (C.c_int32, np.int32),
    "goal_cell": (C.c_int32, np.int32), "episode": (C.c_uint32, np.uint32), "local_map": (C.c_uint8, np.uint8),
    "term_local_map": (C.c_uint8, np.uint8),
    "rel_goal": (C.c_float, np.float32), "velocity": (C.c_float, np.float32), "reward": (C.c_float, np.float32),
    "done": (C.c_uint8, np.uint8), "flags": (C.c_uint8, np.uint8), "term_rel_goal": (C.c_float, np.float32),
    "term_velocity": (C.c_float, np.float32), "fin_return": (C.c_float, np.float32), "fin_length": (C.c_int32, np.int32),
}

INF = 0x7FFFFFFF


def p_threshold(p_occ: float) -> int:
    """SPEC.md §3: min(2^32-1, floor(p_occ * 2^32)) in fp64."""
    return int(min(4294967295, math.floor(float(p_occ) * 4294967296.0)))


def _ptr(a, ty):
    return a.ctypes.data_as(C.POINTER(ty))


def scenario(seed, env_gid, episode, G, p_occ=0.10, goal_mode=0, block_shift=3):
    occ = np.zeros((G, G), np.uint8)
    start = np.zeros(3, np.float32)
    goal = np.zeros(2, np.float32)
    cells = np.zeros(4, np.int32)
    lib().orc_scenario(seed, env_gid, episode, G, p_threshold(p_occ), goal_mode, block_shift, _ptr(occ, C.c_uint8),
                       _ptr(start, C.c_float), _ptr(goal, C.c_float), _ptr(cells, C.c_int32))
    return occ, start, goal, cells


def flow_field(occ, gi, gj):
    """occ u8 [G,G] -> (cost i32, dir u8, flow u8)."""
    occ = np.ascontiguousarray(occ, np.uint8)
    G = occ.shape[0]
    cost = np.zeros((G, G), np.int32)
    d = np.zeros((G, G), np.uint8)
    flow = np.zeros((G, G), np.uint8)
    lib().orc_flow_field(_ptr(occ, C.c_uint8), G, int(gi), int(gj), _ptr(cost, C.c_int32), _ptr(d, C.c_uint8), _ptr(flow, C.c_uint8))
    return cost, d, flow


def crop(flow, W, ci, cj):
    flow = np.ascontiguousarray(flow, np.uint8)
    out = np.zeros((W, W), np.uint8)
    lib().orc_crop(_ptr(flow, C.c_uint8), flow.shape[0], W, int(ci), int(cj), _ptr(out, C.c_uint8))
    return out


def scan(grid_map, pose, beams=360, range_max=3.5, flow_mode=True):
    """SPEC.md §9: (ranges f32[beams], is_collision2 flag) of one pose on a flow image (flow_mode) or an occupancy plane."""
    m = np.ascontiguousarray(grid_map, np.uint8)
    out = np.zeros(beams, np.float32)
    hit = lib().orc_scan(_ptr(m, C.c_uint8), 1 if flow_mode else 0, m.shape[0], np.float32(pose[0]), np.float32(pose[1]),
                         np.float32(pose[2]), beams, np.float32(range_max), _ptr(out, C.c_float))
    return out, int(hit)


def sincos(a):
    s, c = C.c_float(), C.c_float()
    lib().orc_sincos(np.float32(a), C.byref(s), C.byref(c))
    return np.float32(s.value), np.float32(c.value)


class OracleVectorEnv:
    """Sequential CPU restatement of the batched env (SPEC.md §7).  local_map is [N,2,W,W]."""

    def __init__(self, num_envs, grid=128, window=100, max_steps=200, goal_mode=0, p_occ=0.10, seed=0,
                 env_id_base=0, dt=0.1, block_shift=3):
        self.N, self.G, self.W = num_envs, grid, window
        self._L = lib()
        self._h = self._L.orc_env_create(num_envs, grid, window, max_steps, goal_mode, p_threshold(p_occ),
                                         seed, env_id_base, np.float32(dt), block_shift)
        N, G, W = num_envs, grid, window
        shapes = {"occ": (N, G, G), "dir": (N, G, G), "flow": (N, G, G), "cost": (N, G, G), "pose": (N, 3),
                  "goal": (N, 2), "d_first": (N,), "ep_return": (N,), "steps": (N,), "goal_cell": (N, 2),
                  "episode": (N,), "local_map": (N, 2, W, W), "term_local_map": (N, 2, W, W), "rel_goal": (N, 2), "velocity": (N, 2),
                  "reward": (N,), "done": (N,), "flags": (N,), "term_rel_goal": (N, 2), "term_velocity": (N, 2),
                  "fin_return": (N,), "fin_length": (N,)}
        for name, shp in shapes.items():
            p = getattr(self._L, "orc_env_" + name)(self._h)
            setattr(self, name, np.ctypeslib.as_array(p, shape=shp))

    def reset(self, mask=None):
        if mask is None:
            self._L.orc_env_reset(self._h)
        else:
            m = np.ascontiguousarray(mask, np.uint8)
            self._L.orc_env_reset_masked(self._h, _ptr(m, C.c_uint8))
        return self.obs()

    def step(self, actions):
        a = np.ascontiguousarray(actions, np.int64)
        self._L.orc_env_step(self._h, _ptr(a, C.c_int64))
        return self.obs(), self.reward, self.done, self.flags

    def scan(self, beams=360, range_max=3.5):
        """SPEC.md §9 on every env's current scenario and pose: (ranges f32[N,beams], hit u8[N])."""
        out = np.zeros((self.N, beams), np.float32)
        hit = np.zeros(self.N, np.uint8)
        for n in range(self.N):
            out[n], hit[n] = scan(self.flow[n], self.pose[n], beams, range_max, True)
        return out, hit

    def obs(self):
        return {"local_map": self.local_map, "relative_goal": self.rel_goal, "velocity": self.velocity}

    @property
    def error_word(self):
        return self._L.orc_env_error_word(self._h)

    def close(self):
        if self._h:
            self._L.orc_env_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
