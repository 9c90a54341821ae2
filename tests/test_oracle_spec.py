"""CPU: properties of the [SPEC] parts of the oracle (no reference implementation exists for these;
parity is unpinned against the reference and pinned by SPEC.md)."""
import numpy as np
import pytest

import oracle
from maps import special_cases

INF = oracle.INF
DIRS = [(1, 0), (1, 1), (0, 1), (-1, 1), (-1, 0), (-1, -1), (0, -1), (1, -1)]


def check_field(occ, goal, cost, d, flow):
    G = occ.shape[0]
    gi, gj = goal
    ok_goal = 0 <= gi < G and 0 <= gj < G and occ[gi, gj] == 0
    if not ok_goal:
        assert (cost == INF).all() and (d == 8).all()
        return
    assert cost[gi, gj] == 0 and d[gi, gj] == 8
    assert (cost[occ != 0] == INF).all()
    reach = cost != INF
    pad = np.full((G + 2, G + 2), INF, np.int64)
    pad[1:-1, 1:-1] = cost
    nmin = np.minimum(np.minimum(pad[:-2, 1:-1], pad[2:, 1:-1]), np.minimum(pad[1:-1, :-2], pad[1:-1, 2:]))
    inner = reach.copy()
    inner[gi, gj] = False
    assert (nmin[inner] == cost[inner] - 1).all()          # BFS: some 4-neighbour is exactly one closer
    # unreachable free cells have no reachable 4-neighbour
    lost = (occ == 0) & ~reach
    assert (nmin[lost] == INF).all()
    # flow: points to a strictly lower admissible neighbour with the minimum cost; first in scan order
    for (i, j) in np.argwhere(reach)[:: max(1, reach.sum() // 400)]:
        best, bd = cost[i, j], 8
        for k, (di, dj) in enumerate(DIRS):
            ni, nj = i + di, j + dj
            if not (0 <= ni < G and 0 <= nj < G) or occ[ni, nj]:
                continue
            if k & 1 and (occ[ni, j] or occ[i, nj]):
                continue
            if cost[ni, nj] < best:
                best, bd = cost[ni, nj], k
        assert d[i, j] == bd
    assert np.array_equal(flow, np.where(occ != 0, 255, d * 28).astype(np.uint8))


@pytest.mark.parametrize("G", [16, 32, 64, 100, 128])
def test_flow_field_properties(G):
    for name, occ, goal in special_cases(G):
        cost, d, flow = oracle.flow_field(occ, goal[0], goal[1])
        check_field(occ, goal, cost, d, flow)


def test_serpentine_is_deep():
    from maps import serpentine
    occ, g = serpentine(128)
    cost, _, _ = oracle.flow_field(occ, *g)
    assert cost[cost != INF].max() > 4096       # exercises cost bit-planes 8..12 on the GPU


def test_open_map_is_manhattan():
    G = 64
    occ = np.zeros((G, G), np.uint8)
    cost, d, _ = oracle.flow_field(occ, 10, 20)
    ii, jj = np.meshgrid(np.arange(G), np.arange(G), indexing="ij")
    assert np.array_equal(cost, np.abs(ii - 10) + np.abs(jj - 20))
    assert d[12, 22] == 5 and d[12, 20] == 4 and d[8, 18] == 1 and d[10, 25] == 6   # SW, W, NE, S


def test_scenario_properties():
    for G, mode, bs in [(128, 0, 3), (64, 1, 3), (100, 0, 0), (32, 0, 2), (128, 1, 4)]:
        for env in range(6):
            occ, start, goal, cells = oracle.scenario(7, env, env % 3, G, p_occ=0.1, goal_mode=mode, block_shift=bs)
            si, sj, gi, gj = cells
            assert occ[0].all() and occ[-1].all() and occ[:, 0].all() and occ[:, -1].all()
            assert not occ[si - 2:si + 3, sj - 2:sj + 3].any() and not occ[gi - 2:gi + 3, gj - 2:gj + 3].any()
            assert 3 <= si <= G - 4 and 3 <= sj <= G - 4
            if mode == 1:
                assert (gi, gj) == (G - 8, G - 8)
            if G >= 64:
                assert (gi - si) ** 2 + (gj - sj) ** 2 >= 400
            assert start[0] == np.float32(si) * np.float32(0.05) and goal[1] == np.float32(gj) * np.float32(0.05)
            assert -np.pi < start[2] <= np.float32(np.pi)
            again = oracle.scenario(7, env, env % 3, G, p_occ=0.1, goal_mode=mode, block_shift=bs)
            assert np.array_equal(again[0], occ)
    a = oracle.scenario(1, 5, 0, 64)[0]
    assert not np.array_equal(a, oracle.scenario(2, 5, 0, 64)[0])
    assert not np.array_equal(a, oracle.scenario(1, 6, 0, 64)[0])
    assert not np.array_equal(a, oracle.scenario(1, 5, 1, 64)[0])


def test_sharding_invariance():
    """Results depend on the global env id only: two shards of 4 == one batch of 8."""
    full = oracle.OracleVectorEnv(8, grid=64, window=32, seed=11)
    lo = oracle.OracleVectorEnv(4, grid=64, window=32, seed=11, env_id_base=0)
    hi = oracle.OracleVectorEnv(4, grid=64, window=32, seed=11, env_id_base=4)
    for e in (full, lo, hi):
        e.reset()
    rng = np.random.default_rng(3)
    for _ in range(60):
        a = rng.integers(0, 28, 8)
        full.step(a); lo.step(a[:4]); hi.step(a[4:])
        assert np.array_equal(full.reward, np.concatenate([lo.reward, hi.reward]))
        assert np.array_equal(full.local_map, np.concatenate([lo.local_map, hi.local_map]))
        assert np.array_equal(full.done, np.concatenate([lo.done, hi.done]))


def test_episode_semantics():
    env = oracle.OracleVectorEnv(4, grid=64, window=32, max_steps=5, seed=2, p_occ=0.0)
    obs = env.reset()
    assert (obs["velocity"] == 0).all() and np.allclose(obs["relative_goal"][:, 0], env.d_first)
    for t in range(1, 6):
        _, r, done, flags = env.step(np.full(4, 3))             # action 3 = (0, 0): nothing moves
        assert np.allclose(r, -0.05)                            # eps * (d_first - d) = 0, r_s = -0.05
        assert (done == (t == 5)).all() and ((flags & 4 != 0) == (t == 5)).all()
    assert (env.fin_length == 5).all() and np.allclose(env.fin_return, -0.25)
    assert (env.episode == 1).all() and (env.steps == 0).all()
    env.step(np.array([3, 99, -1, 27]))
    assert env.error_word & 1


# ---------------------------------------------------------------------------------------------------
# L: LiDAR scan synthesis (SPEC.md §9)
# ---------------------------------------------------------------------------------------------------
def _room(G):
    occ = np.zeros((G, G), np.uint8)
    occ[0, :] = occ[-1, :] = occ[:, 0] = occ[:, -1] = 1
    return occ


def test_scan_empty_room_axis_beams_are_exact():
    """Robot on a cell centre of an empty walled room: the four axis beams end on the wall faces, (k + 0.5) cells away."""
    G = 64
    occ = _room(G)
    r, hit = oracle.scan(occ, (1.6, 1.0, 0.0), beams=4, range_max=3.5, flow_mode=False)   # cell (32, 20)
    assert hit == 0
    assert r[0] == np.float32(np.float32(30.5) * np.float32(0.05))      # +x: wall row 63 begins at u = 63, u0 = 32.5
    assert r[2] == np.float32(np.float32(31.5) * np.float32(0.05))      # -x: wall row 0 ends at u = 1
    assert abs(r[1] - 42.5 * 0.05) < 1e-4 and abs(r[3] - 19.5 * 0.05) < 1e-4   # +-y (cos(pi/2) is ~-4e-8, not 0)
    # beyond the sensor range nothing returns
    r, _ = oracle.scan(occ, (1.6, 1.0, 0.0), beams=4, range_max=1.0, flow_mode=False)
    assert np.isinf(r[0]) and np.isinf(r[1]) and np.isinf(r[2]) and r[3] == np.float32(np.float32(19.5) * np.float32(0.05))


def test_scan_blocked_robot_cell_and_modes():
    G = 32
    occ = _room(G)
    occ[10, 10] = 1
    r, hit = oracle.scan(occ, (0.5, 0.5, 0.3), beams=16, flow_mode=False)     # robot cell (10, 10) is occupied
    assert (r == 0).all() and hit == 0                                        # zeros are "no reading" (train.py:146)
    r, hit = oracle.scan(occ, (-1.0, 0.5, 0.0), beams=16, flow_mode=False)    # out of the grid
    assert (r == 0).all() and hit == 0
    # a flow image (255 = occupied, other values are direction codes) gives the same scan as its occupancy plane
    cost, d, flow = oracle.flow_field(occ, 20, 20)
    a, ha = oracle.scan(occ, (0.8, 0.7, 1.0), beams=90, flow_mode=False)
    b, hb = oracle.scan(flow, (0.8, 0.7, 1.0), beams=90, flow_mode=True)
    assert np.array_equal(a.view(np.uint32), b.view(np.uint32)) and ha == hb


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_scan_hits_first_occupied_cell_along_the_ray(seed):
    """Independent fp64 check: the ray is free up to the reported range and blocked just beyond it."""
    G = 128
    occ, start, goal, cells = oracle.scenario(seed, 3, 0, G)
    rng = np.random.default_rng(seed)
    free = np.argwhere(occ == 0)
    B = 72
    for (i, j) in free[rng.integers(0, len(free), 12)]:
        x, y, yaw = np.float32(i * 0.05 + rng.uniform(-0.02, 0.02)), np.float32(j * 0.05 + rng.uniform(-0.02, 0.02)), np.float32(rng.uniform(-3.1, 3.1))
        r, hit = oracle.scan(occ, (x, y, yaw), beams=B, range_max=3.5, flow_mode=False)
        inc = 2 * np.pi / B
        for k in range(B):
            th = float(yaw) + k * inc
            c, s = np.cos(th), np.sin(th)
            def blocked(t):
                u, v = float(x) * 20 + 0.5 + t * 20 * c, float(y) * 20 + 0.5 + t * 20 * s
                ii, jj = int(np.floor(u)), int(np.floor(v))
                return not (0 <= ii < G and 0 <= jj < G) or occ[ii, jj] != 0
            if np.isinf(r[k]):
                ts = np.arange(0.0, 3.5 - 2e-3, 0.004)
            else:
                assert 0 < r[k] <= 3.5 + 1e-5
                ts = np.arange(0.0, max(0.0, float(r[k]) - 2e-3), 0.004)
                assert blocked(float(r[k]) + 2e-3), (i, j, k)
            assert not any(blocked(t) for t in ts), (i, j, k)
        # hit == the reference's is_collision2 on the list the trainer would keep (finite, non-zero ranges)
        kept = np.array([v for v in r if np.isfinite(v) and v != 0], np.float64)
        ref = oracle.lib().orc_ref_is_collision2(kept.ctypes.data_as(oracle.C.POINTER(oracle.C.c_double)), len(kept)) if len(kept) else 0
        assert hit == ref
