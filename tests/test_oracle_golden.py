"""CPU: the oracle's restatements of the reference functions against golden vectors produced by
EXECUTING the unmodified reference (oracle/make_golden.py -> tests/golden/ref_golden.json)."""
import ctypes as C
import math

import numpy as np
import pytest

import oracle

L = oracle.lib()
REL = 1e-5  # north_star tolerance for fp32 poses / rewards vs the fp64 reference


def test_action_table(golden):
    for i, (v, w) in enumerate(golden["action_table"]):
        rv, rw = C.c_double(), C.c_double()
        L.orc_ref_action(i, C.byref(rv), C.byref(rw))
        assert (rv.value, rw.value) == (v, w)


def test_footprint_cells(golden):
    ij = np.zeros(2 * 100, np.int32)
    n = L.orc_ref_footprint(100, 0.05, 5.0, 0.13, ij.ctypes.data_as(C.POINTER(C.c_int32)))
    assert n == 21 == len(golden["footprint"])
    assert ij[:2 * n].reshape(-1, 2).tolist() == golden["footprint"]
    # SPEC.md §1: the footprint equals the integer mask di^2 + dj^2 <= 6 around (50, 50)
    mask = sorted([50 + di, 50 + dj] for di in range(-2, 3) for dj in range(-2, 3) if di * di + dj * dj <= 6)
    assert mask == sorted(golden["footprint"])


def test_is_collision(golden):
    for case in golden["is_collision"]:
        m = np.zeros((100, 100), np.int32)
        for (i, j) in case["cells"]:
            m[i, j] = case["value"]
        got = L.orc_ref_is_collision(m.ctypes.data_as(C.POINTER(C.c_int32)), 100, 0.05, 5.0, 0.13)
        assert bool(got) == case["expect"], case
        # [SPEC] collision on a global flow image == reference is_collision on the occupancy crop
        flow = np.where(m > 0, 255, 0).astype(np.uint8)
        assert bool(L.orc_collision(flow.ctypes.data_as(C.POINTER(C.c_uint8)), 100, 50, 50)) == case["expect"]


def test_is_collision2(golden):
    for case in golden["is_collision2"]:
        scan = np.array([np.nan if r is None else r for r in case["scan"]], np.float64)
        got = L.orc_ref_is_collision2(scan.ctypes.data_as(C.POINTER(C.c_double)), len(scan))
        assert bool(got) == case["expect"], case


def test_is_goal_and_done(golden):
    for case in golden["is_goal"]:
        assert bool(L.orc_ref_is_goal(case["d"])) == case["expect"]
    for case in golden["is_done"]:
        assert bool(L.orc_ref_is_done(case["col"], case["goal"])) == case["expect"]


def test_reward_sequences(golden):
    for seq in golden["reward_sequences"]:
        pre = C.c_double(0.0)
        d_first32 = None
        for s in seq:
            got = L.orc_ref_reward(s["d"], s["col"], s["goal"], s["is_first"], C.byref(pre))
            assert got == s["expect"], s                    # fp64 restatement: exact
            # fp32 SPEC formula within the stated tolerance
            d = np.float32(s["d"])
            if s["is_first"]:
                d_first32 = d
            r_g = np.float32(1.0) if s["goal"] else np.float32(0.05) * (d_first32 - d)
            r32 = (r_g + np.float32(-1.0 if s["col"] else 0.0)) + np.float32(-0.05)
            assert abs(float(r32) - s["expect"]) <= REL * max(1.0, abs(s["expect"]))


def test_pi_to_pi(golden):
    for case in golden["pi_to_pi"]:
        assert L.orc_ref_pi_to_pi(case["a"]) == case["expect"]
        got32 = L.orc_pi_to_pi(np.float32(case["a"]))
        # fp32 wrap agrees modulo 2*pi within tolerance (the +-pi boundary itself is a seam)
        diff = (got32 - case["expect"] + math.pi) % (2 * math.pi) - math.pi
        assert abs(diff) <= 1e-5 * max(1.0, abs(case["a"])), case
    assert L.orc_pi_to_pi(np.float32(math.pi)) == pytest.approx(math.pi, abs=1e-6)
    assert L.orc_pi_to_pi(np.float32(-math.pi)) == pytest.approx(math.pi, abs=1e-6)   # (-pi, pi]


def test_relative_goal(golden):
    out = (C.c_double * 2)()
    for case in golden["relative_goal"]:
        (gx, gy), (x, y, yaw) = case["goal"], case["pose"]
        L.orc_ref_relative_goal(gx, gy, x, y, yaw, out)
        assert [out[0], out[1]] == case["expect"]
        # fp32 SPEC path (custom atan2 / sqrt) within tolerance
        f = np.float32
        dx, dy = f(f(gx) - f(x)), f(f(gy) - f(y))
        d32 = L.orc_dist(dx, dy)
        b32 = L.orc_pi_to_pi(f(L.orc_atan2(dy, dx)) - f(yaw))
        assert abs(d32 - case["expect"][0]) <= REL * max(1.0, case["expect"][0])
        diff = (b32 - case["expect"][1] + math.pi) % (2 * math.pi) - math.pi
        assert abs(diff) <= 2e-5, case


def test_velocity_sequences(golden):
    out = (C.c_double * 2)()
    for seq in golden["velocity_sequences"]:
        prev = None
        for s in seq:
            x, y, yaw = s["pose"]
            if s["is_first"]:
                prev = (x, y, yaw)
            L.orc_ref_velocity(x, y, yaw, prev[0], prev[1], prev[2], out)
            assert [out[0], out[1]] == s["expect"]
            prev = (x, y, yaw)


def test_temporal_maps_semantics(golden):
    """a12: on is_first both channels hold the first frame, afterwards [previous, current], oldest first."""
    env = oracle.OracleVectorEnv(1, grid=64, window=32, seed=5)
    env.reset()
    lm = env.local_map.copy()
    assert np.array_equal(lm[0, 0], lm[0, 1])
    ids = [t["channel_ids"] for t in golden["temporal_maps"]]
    assert ids == [[1, 1], [1, 2], [2, 3], [4, 4], [4, 5]]
    prev_new = lm[0, 1].copy()
    for _ in range(5):
        _, _, done, _ = env.step(np.array([24]))
        if done[0]:
            assert np.array_equal(env.local_map[0, 0], env.local_map[0, 1])
        else:
            assert np.array_equal(env.local_map[0, 0], prev_new)
        prev_new = env.local_map[0, 1].copy()


def test_sincos_atan2_accuracy():
    rng = np.random.default_rng(0)
    for a in np.concatenate([rng.uniform(-math.pi, math.pi, 2000), [0.0, math.pi, -math.pi, math.pi / 2]]):
        s, c = oracle.sincos(a)
        assert abs(float(s) - math.sin(np.float32(a))) < 3e-7
        assert abs(float(c) - math.cos(np.float32(a))) < 3e-7
    for _ in range(2000):
        y, x = rng.uniform(-7, 7, 2).astype(np.float32)
        assert abs(L.orc_atan2(y, x) - math.atan2(y, x)) < 1e-6
    assert L.orc_atan2(np.float32(0), np.float32(0)) == 0.0
    assert L.orc_atan2(np.float32(0), np.float32(-1)) == pytest.approx(math.pi, abs=1e-6)


def test_constants(golden):
    c = golden["constants"]
    assert (c["MAP_RANGE"], c["MAP_GRID_NUM"], c["ROBOT_RSIZE"], c["MAP_RESOLUTION"], c["GOAL_THRESHOLHD"]) == \
        (5.0, 100, 0.13, 0.05, 0.5)
