"""Hand-made occupancy maps shared by the CPU and GPU tests."""
import numpy as np


def bordered(G):
    occ = np.zeros((G, G), np.uint8)
    occ[0, :] = occ[-1, :] = occ[:, 0] = occ[:, -1] = 1
    return occ


def serpentine(G, gap=1):
    """Walls on every other row with alternating gaps: BFS depth ~ G*G/2 (exercises cost bits >= 8)."""
    occ = bordered(G)
    for k, i in enumerate(range(2, G - 2, 2)):
        occ[i, 1:G - 1] = 1
        if k % 2 == 0:
            occ[i, G - 1 - gap:G - 1] = 0
        else:
            occ[i, 1:1 + gap] = 0
    return occ, (1, 1)


def rooms(G, seed=0):
    rng = np.random.default_rng(seed)
    occ = bordered(G)
    for _ in range(G // 8):
        i = int(rng.integers(2, G - 2))
        occ[i, 1:G - 1] = 1
        occ[i, int(rng.integers(1, G - 1))] = 0
        j = int(rng.integers(2, G - 2))
        occ[1:G - 1, j] = 1
        occ[int(rng.integers(1, G - 1)), j] = 0
    free = np.argwhere(occ == 0)
    g = free[int(rng.integers(0, len(free)))]
    return occ, (int(g[0]), int(g[1]))


def noise(G, p, seed=0, values=(1,)):
    rng = np.random.default_rng(seed)
    occ = (rng.random((G, G)) < p).astype(np.uint8) * np.uint8(rng.choice(values))
    free = np.argwhere(occ == 0)
    g = free[int(rng.integers(0, len(free)))]
    return occ, (int(g[0]), int(g[1]))


def special_cases(G):
    cases = []
    cases.append(("open", bordered(G), (G // 2, G // 3)))
    cases.append(("no_border_open", np.zeros((G, G), np.uint8), (0, 0)))
    cases.append(("no_border_corner", np.zeros((G, G), np.uint8), (G - 1, G - 1)))
    full = np.ones((G, G), np.uint8)
    cases.append(("all_blocked", full, (G // 2, G // 2)))
    occ = bordered(G)
    occ[G // 2, G // 2] = 1
    cases.append(("goal_blocked", occ, (G // 2, G // 2)))
    cases.append(("goal_out_of_grid", bordered(G), (-1, 5)))
    cases.append(("goal_out_of_grid_hi", bordered(G), (3, G)))
    occ = bordered(G)
    occ[3:6, 3] = occ[3:6, 7] = occ[3, 3:8] = occ[5, 3:8] = 1   # sealed pocket: unreachable free cell
    occ[4, 4:7] = 0
    cases.append(("sealed_pocket", occ, (G - 3, G - 3)))
    s, g = serpentine(G)
    cases.append(("serpentine", s, g))
    s2, g2 = serpentine(G, gap=2)
    cases.append(("serpentine_gap2", s2, (G - 2, G - 2)))
    for sd in range(3):
        o, g = rooms(G, sd)
        cases.append((f"rooms{sd}", o, g))
    for sd, p in enumerate((0.05, 0.2, 0.35, 0.45)):
        o, g = noise(G, p, sd, values=(1, 255, 7))
        cases.append((f"noise{p}", o, g))
    return cases
