"""`FFMP` — single-env compatibility object with the reference's method surface
(/root/reference/src/gym_ffmp/envs/ffmp.py:22-188), evaluated by the CUDA operators.

train.py's only env call is `env.rewarder2(scan, relative_goal, is_first)` (train.py:577); `rewarder`
is the map-based variant the batched env reproduces.  Inputs are host scalars / arrays as in the
reference; they are staged to the device, evaluated by `ffmp_op_rewarder`, and read back.
"""
import numpy as np
import torch

from . import ops
from .robot import RobotAction
from .vector_env import (GOAL_THRESHOLD, MAP_GRID_NUM, MAP_RANGE, MAP_RESOLUTION, ROBOT_RSIZE, FFMPVectorEnv,
                         make_spaces)


class FFMP:
    def __init__(self, device="cuda:0", grid=MAP_GRID_NUM, seed=0, **env_kwargs):
        self.device = torch.device(device)
        self._env_kwargs = dict(grid=grid, window=MAP_GRID_NUM, seed=seed, **env_kwargs)
        self._env = None                                                          # 1-env FFMPVectorEnv behind reset() / step()
        self.action = RobotAction()
        self.action_space, self.observation_space, self.state_space = make_spaces(MAP_GRID_NUM)
        self.map_range, self.map_grid_num, self.map_grid_size = MAP_RANGE, MAP_GRID_NUM, MAP_RESOLUTION
        self.map_channels = 1
        self.robot_rsize = ROBOT_RSIZE
        self._d_first = torch.zeros(1, dtype=torch.float32, device=self.device)   # ffmp.py:139 module global, per object here

    def _t(self, x, dtype):
        return torch.as_tensor(np.asarray(x), device=self.device).to(dtype)

    def _map(self, local_map):
        return self._t(local_map, torch.int32).reshape(1, self.map_grid_num, self.map_grid_num)

    def _goal(self, rel_goal):
        return torch.tensor([[float(rel_goal[0]), float(rel_goal[1])]], dtype=torch.float32, device=self.device)

    def _first(self, is_first):
        return torch.tensor([1 if is_first else 0], dtype=torch.uint8, device=self.device)

    @staticmethod
    def _scan(scan_data):
        vals = [float("nan") if r is None else float(r) for r in scan_data] or [float("nan")]
        return torch.tensor([vals], dtype=torch.float64)

    def is_collision(self, local_map_info):                                   # ffmp.py:85-105
        scratch = self._d_first.clone()
        _, _, flags = ops.rewarder(self._map(local_map_info), self._goal((1e9, 0.0)), self._first(False), scratch)
        return bool(int(flags.item()) & 1)

    def is_collision2(self, scan_data):                                        # ffmp.py:108-117
        scratch = self._d_first.clone()
        _, _, flags = ops.rewarder2(self._scan(scan_data).to(self.device), self._goal((1e9, 0.0)), self._first(False), scratch)
        return bool(int(flags.item()) & 1)

    def is_goal(self, cur_relative_goal_dist):                                # ffmp.py:120-127
        scratch = self._d_first.clone()
        lm = torch.zeros((1, self.map_grid_num, self.map_grid_num), dtype=torch.int32, device=self.device)
        _, _, flags = ops.rewarder(lm, self._goal((cur_relative_goal_dist, 0.0)), self._first(False), scratch)
        return bool(int(flags.item()) & 2)

    def is_done(self, is_collision, is_goal):                                 # ffmp.py:160-164
        return bool(is_collision or is_goal)

    def reward_calculator(self, relative_goal_info, is_collision, is_goal, is_first):   # ffmp.py:130-157
        col = torch.tensor([bool(is_collision)], device=self.device)
        goal = torch.tensor([bool(is_goal)], device=self.device)
        r, _, _ = ops.reward_calculator(self._goal(relative_goal_info), col, goal, self._first(is_first), self._d_first)
        return float(r.item())

    def rewarder(self, local_map_info, relative_goal_info, is_first):         # ffmp.py:167-176
        r, d, _ = ops.rewarder(self._map(local_map_info), self._goal(relative_goal_info), self._first(is_first), self._d_first)
        return float(r.item()), bool(d.item())

    def rewarder2(self, scan_data, relative_goal_info, is_first):             # ffmp.py:179-188
        r, d, f = ops.rewarder2(self._scan(scan_data).to(self.device), self._goal(relative_goal_info),
                                self._first(is_first), self._d_first)
        return float(r.item()), bool(d.item()), bool(int(f.item()) & 2)


    # ---- gym (Dec 2020, docker/Dockerfile:55) episode API: reset() -> obs, step(a) -> (obs, reward, done, info) ----------
    # The reference never defined these (ffmp.py:77-83 is commented out; the episode lived in ROS nodes); here
    # they drive one environment of the batched CUDA env (BASELINE config 1: one env on the reference's 100 x 100 map).
    def _single(self):
        if self._env is None:
            self._env = FFMPVectorEnv(1, device=str(self.device), **self._env_kwargs)
            self._act = torch.zeros(1, dtype=torch.int64).pin_memory()
        return self._env

    def _np_obs(self, obs):
        lm = obs["local_map"][0].cpu().numpy()                                    # u8 [2, W, W]: older, newest
        return {"local_map": lm[1].astype(np.int32)[:, :, None],                  # Box(0, 255, (100, 100, 1), int32), ffmp.py:40-58
                "relative_goal": np.array(obs["relative_goal"][0].cpu().numpy(), dtype=np.float32),
                "velocity": np.array(obs["velocity"][0].cpu().numpy(), dtype=np.float32)}, lm

    def seed(self, seed=None):
        if seed is not None:
            self._env_kwargs["seed"] = int(seed)
            self.close()
        return [self._env_kwargs["seed"]]

    def reset(self):
        return self._np_obs(self._single().reset())[0]

    def step(self, action):
        """action: one of the 28 discrete ids the reference's agent emits (train.py:70,345,438)."""
        env = self._single()
        self._act[0] = int(action)
        obs, reward, done, info = env.step_host(self._act)
        o, stack = self._np_obs(obs)
        flags = int(info["flags"][0])
        return o, float(reward[0]), bool(done[0]), {"is_collision": bool(flags & 1), "is_goal": bool(flags & 2),
                                                    "truncated": bool(flags & 4), "local_map_stack": stack}

    def close(self):
        if self._env is not None:
            self._env.close()
            self._env = None


_REGISTRY = {}


def register(id, entry_point=None, **kwargs):
    _REGISTRY[id] = (entry_point, kwargs)


def make(id, **kwargs):
    """make('FFMP-v0') -> FFMP compat object (as gym.make in train.py:456);
    make('FFMPVector-v0', num_envs=...) -> FFMPVectorEnv."""
    if id not in _REGISTRY:
        raise KeyError(f"unknown env id {id!r}")
    entry, defaults = _REGISTRY[id]
    return entry(**{**defaults, **kwargs})


register("FFMP-v0", FFMP)
register("FFMPVector-v0", FFMPVectorEnv)
