"""Generate tests/golden/ref_golden.json by EXECUTING the unmodified reference.

Runs only in the build container (needs /root/reference); the GPU box and the tests read the
committed JSON.  The reference package `gym_ffmp` is imported as-is under the stand-in `gym`
in oracle/gym_shim; the pure helper functions of src/train.py (which cannot be imported: it
needs rospy/kornia/tf) are exec'd from their exact line ranges, unmodified.

    python oracle/make_golden.py
"""
import contextlib
import io
import json
import math
import os
import random
import sys
import textwrap

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/src"
OUT = os.path.join(HERE, "..", "tests", "golden", "ref_golden.json")

sys.dont_write_bytecode = True
sys.path.insert(0, os.path.join(HERE, "gym_shim"))
sys.path.insert(0, REF)

import gym  # noqa: E402  (the shim)
import gym_ffmp  # noqa: E402,F401  (the unmodified reference; registers FFMP-v0)
from gym_ffmp.envs import ffmp as ref_ffmp  # noqa: E402
from gym_ffmp.envs.robot.config import RobotAction  # noqa: E402


def exec_lines(path, first, last, env):
    with open(path) as f:
        lines = f.readlines()[first - 1:last]
    src = textwrap.dedent("".join(lines))
    exec(compile(src, f"{path}:{first}-{last}", "exec"), env)
    return env


def main():
    rng = random.Random(20261018)
    g = {}
    env = gym.make("FFMP-v0")

    # a1: action table (robot/config.py:25-58)
    act = RobotAction()
    g["action_table"] = [[act.commander(i).linear_v, act.commander(i).angular_v] for i in range(28)]

    # a2: spaces (ffmp.py:28-64)
    obs = env.observation_space
    g["spaces"] = {
        "action": {"low": env.action_space.low.tolist(), "high": env.action_space.high.tolist(),
                   "dtype": str(env.action_space.dtype), "shape": list(env.action_space.shape)},
        "local_map": {"shape": list(obs["local_map"].shape), "dtype": str(obs["local_map"].dtype),
                      "low": int(obs["local_map"].low.min()), "high": int(obs["local_map"].high.max())},
        "relative_goal": {"low": obs["relative_goal"].low.tolist(), "high": obs["relative_goal"].high.tolist(),
                          "dtype": str(obs["relative_goal"].dtype)},
        "velocity": {"low": obs["velocity"].low.tolist(), "high": obs["velocity"].high.tolist(),
                     "dtype": str(obs["velocity"].dtype)},
        "state_keys": sorted(env.state_space.spaces.keys()),
    }

    # a3: footprint + is_collision (ffmp.py:85-105)
    empty = np.zeros((100, 100), dtype=np.int64)
    env.is_collision(empty)
    g["footprint"] = [[int(a[0]), int(a[1])] for a in env.robot_grids]
    cases = []
    fixed = [[(50, 50)], [(48, 50)], [(47, 50)], [(48, 48)], [(52, 51)], [(52, 52)], [(49, 48)], [(53, 50)], []]
    for cells in fixed:
        m = empty.copy()
        for (i, j) in cells:
            m[i, j] = 255
        cases.append({"cells": [list(c) for c in cells], "value": 255, "expect": bool(env.is_collision(m))})
    for _ in range(40):
        n = rng.randint(1, 6)
        cells = [(rng.randint(44, 56), rng.randint(44, 56)) for _ in range(n)]
        val = rng.choice([1, 28, 224, 255])
        m = empty.copy()
        for (i, j) in cells:
            m[i, j] = val
        cases.append({"cells": [list(c) for c in cells], "value": val, "expect": bool(env.is_collision(m))})
    g["is_collision"] = cases

    # a4: is_collision2 (ffmp.py:108-117); None encoded as null
    scans = [[None, 0.5, 0.12], [None, 0.5, 0.13], [0.0, 0.2], [], [3.0, 0.1299999], [None], [0.131, 0.13, 0.5]]
    with contextlib.redirect_stdout(io.StringIO()):
        g["is_collision2"] = [{"scan": s, "expect": bool(env.is_collision2(s))} for s in scans]

    # a5: is_goal (ffmp.py:120-127)
    ds = [0.49, 0.5, 0.4999999, 0.5000001, 0.0, 3.0]
    g["is_goal"] = [{"d": d, "expect": bool(env.is_goal(d))} for d in ds]

    # a6: reward_calculator sequences (ffmp.py:130-157), incl. the latched module global
    seqs = []
    base = [(True, 3.0, False, False), (False, 2.9, False, False), (False, 2.8, False, False),
            (False, 2.8, True, False), (False, 0.4, False, True), (False, 0.4, True, True)]
    seq = []
    for (first, d, col, goal) in base:
        seq.append({"is_first": first, "d": d, "col": col, "goal": goal,
                    "expect": env.reward_calculator([d, 0.0], col, goal, first)})
    seqs.append(seq)
    for _ in range(6):
        seq = []
        d = rng.uniform(1.0, 7.0)
        for t in range(12):
            first = t == 0
            d = max(0.0, d + rng.uniform(-0.08, 0.06))
            col = rng.random() < 0.1
            goal = d < 0.5
            seq.append({"is_first": first, "d": d, "col": col, "goal": goal,
                        "expect": env.reward_calculator([d, 0.0], col, goal, first)})
        seqs.append(seq)
    g["reward_sequences"] = seqs

    # a7: is_done (ffmp.py:160-164)
    g["is_done"] = [{"col": c, "goal": gl, "expect": bool(env.is_done(c, gl))}
                    for c in (False, True) for gl in (False, True)]

    # a8: rewarder / rewarder2 (ffmp.py:167-188)
    with contextlib.redirect_stdout(io.StringIO()):
        r2 = env.rewarder2([None, 0.5, 0.12], [2.0, 0.0], True)
    g["rewarder2"] = [{"scan": [None, 0.5, 0.12], "rel_goal": [2.0, 0.0], "is_first": True,
                       "expect": [r2[0], bool(r2[1]), bool(r2[2])]}]
    rw = []
    for cells, relg, first in [([(50, 50)], [2.0, 0.1], True), ([], [2.0, 0.1], True), ([], [1.5, 0.1], False),
                               ([(47, 50)], [0.3, 0.0], False), ([(48, 50)], [0.3, 0.0], False)]:
        m = empty.copy()
        for (i, j) in cells:
            m[i, j] = 255
        r, d = env.rewarder(m, relg, first)
        rw.append({"cells": [list(c) for c in cells], "rel_goal": relg, "is_first": first, "expect": [r, bool(d)]})
    g["rewarder"] = rw

    # a9-a11: train.py:167-188 exec'd verbatim inside a stand-in ROSNode
    ns = {"math": math, "np": np, "copy": __import__("copy")}
    src_env = exec_lines(os.path.join(REF, "train.py"), 167, 188, dict(ns))

    class Pose:
        def __init__(self, x, y, yaw=0.0):
            self.x, self.y, self.yaw = x, y, yaw

    class Node:
        pi_to_pi = src_env["pi_to_pi"]
        relative_goal_calculator = src_env["relative_goal_calculator"]
        robot_velocity_calculator = src_env["robot_velocity_calculator"]

        def __init__(self):
            self.pre_robot_pose = Pose(0.0, 0.0, 0.0)

            class G:
                position = Pose(0.0, 0.0)
            self.global_goal = G()

    node = Node()
    angles = [math.pi, -math.pi, 0.0, 3.2, -3.2, 6.5, -6.5, 1.0, 9.5, -12.0, math.pi - 1e-9, -math.pi + 1e-9]
    angles += [rng.uniform(-10, 10) for _ in range(30)]
    g["pi_to_pi"] = [{"a": a, "expect": node.pi_to_pi(a)} for a in angles]

    rg = []
    for _ in range(60):
        gx, gy = rng.uniform(0.2, 6.2), rng.uniform(0.2, 6.2)
        x, y, yaw = rng.uniform(0.2, 6.2), rng.uniform(0.2, 6.2), rng.uniform(-math.pi, math.pi)
        node.global_goal.position.x, node.global_goal.position.y = gx, gy
        out = node.relative_goal_calculator(Pose(x, y, yaw))
        rg.append({"goal": [gx, gy], "pose": [x, y, yaw], "expect": [float(out[0]), float(out[1])]})
    g["relative_goal"] = rg

    vel = []
    for _ in range(4):
        seq = []
        x, y, yaw = rng.uniform(1, 5), rng.uniform(1, 5), rng.uniform(-3, 3)
        for t in range(10):
            if t:
                x += rng.uniform(-0.06, 0.06)
                y += rng.uniform(-0.06, 0.06)
                yaw = node.pi_to_pi(yaw + rng.uniform(-0.06, 0.06) + (3.0 if t == 5 else 0.0))
            out = node.robot_velocity_calculator(Pose(x, y, yaw), t == 0)
            seq.append({"pose": [x, y, yaw], "is_first": t == 0, "expect": [float(out[0]), float(out[1])]})
        vel.append(seq)
    g["velocity_sequences"] = vel

    # a12: make_temporal_maps (train.py:474-486) exec'd verbatim
    import torch
    tm_env = exec_lines(os.path.join(REF, "train.py"), 474, 486, {"torch": torch, "INPUT_CHANNELS": 2})

    class Holder:
        make_temporal_maps = tm_env["make_temporal_maps"]

        def __init__(self):
            self.map_memory = []

    h = Holder()
    tm = []
    with contextlib.redirect_stdout(io.StringIO()):
        for t, first in enumerate([True, False, False, True, False]):
            frame = torch.full((1, 2, 2), float(t + 1))
            out = h.make_temporal_maps(frame, first)
            tm.append({"frame_id": t + 1, "is_first": first, "shape": list(out.shape),
                       "channel_ids": [int(out[0, 0, 0]), int(out[1, 0, 0])]})
    g["temporal_maps"] = tm

    # constants (ffmp.py:14-19, train.py:49-60,69-70)
    g["constants"] = {"MAP_RANGE": ref_ffmp.MAP_RANGE, "MAP_GRID_NUM": ref_ffmp.MAP_GRID_NUM,
                      "ROBOT_RSIZE": ref_ffmp.ROBOT_RSIZE, "MAP_RESOLUTION": ref_ffmp.MAP_RESOLUTION,
                      "GOAL_THRESHOLHD": ref_ffmp.GOAL_THRESHOLHD, "MAX_STEPS": 200, "NUM_ACTIONS": 28,
                      "INPUT_CHANNELS": 2}

    with open(OUT, "w") as f:
        json.dump(g, f, indent=1, sort_keys=True)
    print("wrote", os.path.normpath(OUT), os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
