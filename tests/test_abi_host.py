"""CPU: the C-ABI library loads and exports every symbol include/ffmp_b200.h declares; host-side
logic (spaces, action table, config validation, loud failure without a GPU).  No compute calls."""
import ctypes as C
import os
import re

import numpy as np
import pytest
import torch

import flow_field_based_motion_planner_b200 as ffmp
from flow_field_based_motion_planner_b200 import native

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    text = open(os.path.join(ROOT, "include", "ffmp_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ffmp_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    L = native.lib()
    syms = header_symbols()
    assert len(syms) >= 18
    for s in syms:
        assert hasattr(L, s), f"{s} declared in include/ffmp_b200.h but not exported"
    assert sorted(native.EXPORTS) == syms
    assert L.ffmp_abi_version() == native.ABI_VERSION


def test_struct_layouts_match_header():
    assert C.sizeof(native.Cfg) == 64
    assert native.Cfg.seed.offset == 48 and native.Cfg.dt.offset == 56
    assert C.sizeof(native.Sizes) == 9 * C.sizeof(C.c_size_t)
    assert C.sizeof(native.Buffers) == 15 * C.sizeof(C.c_void_p)


def test_cfg_validation_without_gpu():
    L = native.lib()
    sz = native.Sizes()
    good = dict(abi_version=1, device=0, num_envs=8, grid=128, window=100, ring=8, slots=3, max_steps=200,
                goal_mode=0, block_shift=3, p_thresh=1, env_id_base=0, seed=0, dt=0.1, regen_batch=0)
    assert L.ffmp_query_sizes(C.byref(native.Cfg(**good)), C.byref(sz)) == 0
    assert sz.flow == 3 * 8 * 128 * 128 and sz.cost == 4 * sz.flow and sz.frames == 8 * 8 * 100 * 100
    assert sz.state == 8 * 64 and sz.scen == 3 * 8 * 32 and sz.workspace > 0
    for key, bad in [("abi_version", 2), ("num_envs", 0), ("grid", 130), ("grid", 8), ("window", 102), ("ring", 1),
                     ("slots", 1), ("regen_batch", 3), ("regen_batch", 9), ("max_steps", 0), ("goal_mode", 2), ("dt", 0.0), ("grid", 1024), ("grid", 144)]:
        rc = L.ffmp_query_sizes(C.byref(native.Cfg(**{**good, key: bad})), C.byref(sz))
        assert rc < 0, key
        assert len(L.ffmp_last_error()) > 0


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_fails_loudly_without_gpu():
    with pytest.raises(native.NativeError):
        ffmp.FFMPVectorEnv(4)
    L = native.lib()
    h = C.c_void_p()
    cfg = native.Cfg(abi_version=1, device=0, num_envs=8, grid=128, window=100, ring=8, slots=3, max_steps=200,
                     goal_mode=0, block_shift=3, p_thresh=1, env_id_base=0, seed=0, dt=0.1, regen_batch=0)
    assert L.ffmp_create(C.byref(cfg), C.byref(h)) == -2          # FFMP_ERR_DEVICE: no CPU fallback
    assert b"no CUDA device" in L.ffmp_last_error()
    with pytest.raises(native.NativeError):
        ffmp.ops.flow_field(torch.zeros((1, 32, 32), dtype=torch.uint8), torch.zeros((1, 2)))


def test_spaces_match_reference(golden):
    a, o, s = ffmp.make_spaces()
    g = golden["spaces"]
    assert a.low.tolist() == g["action"]["low"] and a.high.tolist() == g["action"]["high"]
    assert str(a.dtype) == g["action"]["dtype"] and list(a.shape) == g["action"]["shape"]
    lm = o["local_map"]
    assert list(lm.shape) == g["local_map"]["shape"] and str(lm.dtype) == g["local_map"]["dtype"]
    assert int(lm.low.min()) == g["local_map"]["low"] and int(lm.high.max()) == g["local_map"]["high"]
    for k in ("relative_goal", "velocity"):
        assert o[k].low.tolist() == g[k]["low"] and o[k].high.tolist() == g[k]["high"] and str(o[k].dtype) == g[k]["dtype"]
    assert sorted(s.keys()) == g["state_keys"]


def test_action_table_matches_reference(golden):
    act = ffmp.RobotAction()
    assert [[c.linear_v, c.angular_v] for c in act.cmd] == golden["action_table"]
    assert ffmp.NUM_ACTIONS == 28
    assert (act.commander(3).linear_v, act.commander(3).angular_v) == (0.0, 0.0)   # first-tick default, train.py:512


def test_p_threshold():
    import oracle
    for p in (0.0, 0.1, 0.3, 0.999999, 1.0):
        assert ffmp.p_threshold(p) == oracle.p_threshold(p)
    assert ffmp.p_threshold(1.0) == 2 ** 32 - 1 and ffmp.p_threshold(0.0) == 0


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "flow_field_based_motion_planner_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in text and "from oracle" not in text and "ffmp_oracle" not in text, f


def test_gym_ffmp_import_shim_mirrors_reference_module_paths():
    """`import gym_ffmp` + the module paths train.py:35-37 uses resolve to the CUDA-backed objects (no compute here)."""
    import importlib
    import sys
    if ROOT not in sys.path:
        sys.path.insert(0, ROOT)
    gym_ffmp = importlib.import_module("gym_ffmp")
    cfg = importlib.import_module("gym_ffmp.envs.robot.config")
    env_mod = importlib.import_module("gym_ffmp.envs.ffmp")
    assert env_mod.FFMP is ffmp.FFMP and importlib.import_module("gym_ffmp.envs").FFMP is ffmp.FFMP
    assert cfg.RobotAction is ffmp.RobotAction and cfg.RobotPose is ffmp.RobotPose
    assert cfg.RobotVelocity is ffmp.RobotVelocity and cfg.RobotState is ffmp.RobotState
    assert gym_ffmp.ENV_ID == "FFMP-v0" and callable(gym_ffmp.make)
    assert (env_mod.MAP_GRID_NUM, env_mod.MAP_RESOLUTION, env_mod.GOAL_THRESHOLD) == (100, 0.05, 0.5)      # ffmp.py:14-19
