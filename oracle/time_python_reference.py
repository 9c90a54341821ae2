"""Time the reference's OWN Python path for BASELINE.json config 1 and commit the result.

Runs only in the build container (needs /root/reference); bench.py on the GPU box reads the committed
tests/golden/python_reference_timing.json and reports it as cpu_baseline.python_reference.

One env on the reference's 100 x 100 map (ffmp.py:15), p_occ = 0.10, seed 0, 1000 uniform random action ids.  Every control
tick runs, UNMODIFIED and in the order of train.py:535-608:
    RobotAction.commander (robot/config.py:57-58), relative_goal_calculator / robot_velocity_calculator / pi_to_pi
    (train.py:167-188, exec'd), make_temporal_maps (train.py:474-486, exec'd), FFMP.rewarder (ffmp.py:167-176: the map-based
    collision test + reward + done).
The parts the reference left to ROS nodes (scenario, flow field, kinematics, crop) come from the C oracle
port, stepped in lock-step; its time is reported separately ("restated").

    python oracle/time_python_reference.py
"""
import contextlib
import io
import json
import math
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = "/root/reference/src"
OUT = os.path.join(ROOT, "tests", "golden", "python_reference_timing.json")
sys.dont_write_bytecode = True
sys.path.insert(0, os.path.join(HERE, "gym_shim"))
sys.path.insert(0, REF)
sys.path.append(ROOT)                       # after the reference: `gym_ffmp` must be the unmodified package, not the shim

import gym  # noqa: E402  (the stand-in)
import gym_ffmp  # noqa: E402,F401  (the unmodified reference)
from gym_ffmp.envs.robot.config import RobotAction, RobotPose  # noqa: E402
from make_golden import exec_lines  # noqa: E402
import oracle  # noqa: E402


def main(steps=1000):
    import torch
    env = gym.make("FFMP-v0")
    act = RobotAction()
    ns = exec_lines(os.path.join(REF, "train.py"), 167, 188, {"math": math, "np": np, "copy": __import__("copy")})
    tm = exec_lines(os.path.join(REF, "train.py"), 474, 486, {"torch": torch, "INPUT_CHANNELS": 2})

    class Node:
        pi_to_pi = ns["pi_to_pi"]
        relative_goal_calculator = ns["relative_goal_calculator"]
        robot_velocity_calculator = ns["robot_velocity_calculator"]
        make_temporal_maps = tm["make_temporal_maps"]

        def __init__(self):
            self.map_memory = []
            self.pre_robot_pose = RobotPose(0.0, 0.0, 0.0)

            class G:
                position = RobotPose(0.0, 0.0, 0.0)
            self.global_goal = G()

    node = Node()
    sim = oracle.OracleVectorEnv(1, grid=100, window=100, p_occ=0.10, seed=0)
    rng = np.random.default_rng(0)
    actions = rng.integers(0, 28, steps)
    t_ref = t_sim = 0.0
    t0 = time.perf_counter()
    sim.reset()
    t_sim += time.perf_counter() - t0
    is_first, n_done, ret = True, 0, 0.0
    sink = io.StringIO()
    with contextlib.redirect_stdout(sink):          # make_temporal_maps and is_collision2 print every call
        for t in range(steps):
            a = int(actions[t])
            # --- reference code on the current observation (train.py:535-545, 577) ---
            t0 = time.perf_counter()
            cmd = act.commander(a)
            pose = RobotPose(float(sim.pose[0, 0]), float(sim.pose[0, 1]), float(sim.pose[0, 2]))
            node.global_goal.position.x, node.global_goal.position.y = float(sim.goal[0, 0]), float(sim.goal[0, 1])
            rel = node.relative_goal_calculator(pose)
            vel = node.robot_velocity_calculator(pose, is_first)
            frame = torch.from_numpy(sim.local_map[0, 1].astype(np.float32))[None]
            maps = node.make_temporal_maps(frame, is_first)
            occ_map = (sim.local_map[0, 1] == 255).astype(np.int64)       # the occupancy the flow image carries
            reward, done = env.rewarder(occ_map, rel, is_first)
            t_ref += time.perf_counter() - t0
            ret += reward
            is_first = False
            # --- the world the reference left to ROS: kinematics, crop, and on an episode end scenario + flow field ---
            t0 = time.perf_counter()
            sim.step(np.array([a]))
            t_sim += time.perf_counter() - t0
            if sim.done[0]:
                is_first = True
                n_done += 1
            del cmd, vel, maps
    total = t_ref + t_sim
    out = {"workload": "BASELINE configs[0]: 1 env, 100x100 grid, p_occ 0.10, seed 0, 1000 uniform random action ids",
           "steps": steps, "episodes_finished": n_done, "env_steps_per_s": steps / total, "seconds": total,
           "reference_code_seconds": t_ref, "restated_code_seconds": t_sim, "reference_code_fraction": t_ref / total,
           "cores": 1, "where": "build container (8 vCPU), CPython %d.%d; not measured on the GPU box: /root/reference does not travel" % sys.version_info[:2],
           "reference_code": "RobotAction.commander robot/config.py:57-58; relative_goal_calculator, robot_velocity_calculator, "
                             "pi_to_pi train.py:167-188; make_temporal_maps train.py:474-486; FFMP.rewarder ffmp.py:167-176",
           "restated_code": "oracle/ffmp_oracle.c: scenario, flow field, kinematics, crop (SPEC.md), called through ctypes"}
    with open(OUT, "w") as f:
        json.dump(out, f, indent=1)
    sys.stdout.write(json.dumps(out, indent=1) + "\n")


if __name__ == "__main__":
    main()
