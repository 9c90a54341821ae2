"""Stateless CUDA operators behind the C-ABI: what the reference's external ROS nodes computed.

  generate_scenarios -> replaces the episode_manager node   (/root/reference/src/train.py:86-90,128-132)
  flow_field         -> replaces the /bev/* flow-image node  (/root/reference/src/train.py:84,116-121)
  rewarder           -> batched FFMP.rewarder                (/root/reference/src/gym_ffmp/envs/ffmp.py:167-176)
  scan               -> replaces Gazebo's /scan LaserScan    (/root/reference/src/train.py:87,144-150; SPEC.md §9)
No CPU fallback: every function raises if the CUDA library or a B200 is missing.
"""
import ctypes as C

import torch

from . import native
from .vector_env import p_threshold


def _dev_index(t: torch.Tensor) -> int:
    if t.device.type != "cuda":
        raise native.NativeError("operators need CUDA tensors; there is no CPU fallback")
    return t.device.index if t.device.index is not None else torch.cuda.current_device()


def _stream(device):
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def generate_scenarios(env_gids, episodes, grid, p_occ=0.10, goal_mode=0, block_shift=3, seed=0):
    """env_gids, episodes: integer tensors [n] on the device -> (occ u8[n,G,G], scen i32[n,8])."""
    L = native.lib()
    dev = _dev_index(env_gids)
    n = env_gids.numel()
    gids = env_gids.to(torch.int32).contiguous()      # bit pattern of u32
    eps = episodes.to(torch.int32).contiguous()
    occ = torch.empty((n, grid, grid), dtype=torch.uint8, device=env_gids.device)
    scen = torch.empty((n, 8), dtype=torch.int32, device=env_gids.device)
    with torch.cuda.device(env_gids.device):
        native.check(L.ffmp_op_scenarios(dev, n, grid, p_threshold(p_occ), goal_mode, block_shift, seed,
                                         C.c_void_p(gids.data_ptr()), C.c_void_p(eps.data_ptr()),
                                         C.c_void_p(occ.data_ptr()), C.c_void_p(scen.data_ptr()),
                                         _stream(env_gids.device)), "ffmp_op_scenarios")
    return occ, scen


def flow_field(occ, goal_cells, want_cost=True, out=None, workspace=None):
    """occ u8[n,G,G] (non-zero = occupied), goal_cells int[n,2] -> (cost i32[n,G,G] | None, flow u8[n,G,G]).

    out = (cost, flow) and workspace (u8, flow_field_workspace(n, G) bytes) may be passed to reuse buffers between calls."""
    L = native.lib()
    dev = _dev_index(occ)
    assert occ.dtype == torch.uint8 and occ.dim() == 3 and occ.shape[1] == occ.shape[2]
    occ = occ.contiguous()
    n, G = occ.shape[0], occ.shape[1]
    goals = goal_cells
    if goals.dtype != torch.int32 or goals.device != occ.device or not goals.is_contiguous():
        goals = goal_cells.to(device=occ.device, dtype=torch.int32).contiguous()
    need_cost = want_cost
    if out is not None:
        cost, flow = out
        assert flow.dtype == torch.uint8 and flow.shape == occ.shape and flow.is_contiguous() and flow.device == occ.device
        if need_cost:
            assert cost.dtype == torch.int32 and cost.shape == occ.shape and cost.is_contiguous() and cost.device == occ.device
    else:
        cost = torch.empty((n, G, G), dtype=torch.int32, device=occ.device) if need_cost else None
        flow = torch.empty((n, G, G), dtype=torch.uint8, device=occ.device)
    if n == 0:
        return cost, flow
    ws_bytes = L.ffmp_op_flow_field_workspace(n, G)
    if n > 0 and ws_bytes == 0:
        raise native.NativeError(f"flow_field: grid {G} is not supported by this build")
    ws = workspace
    if ws is None:
        ws = torch.empty((max(ws_bytes, 16),), dtype=torch.uint8, device=occ.device)
    assert ws.dtype == torch.uint8 and ws.numel() >= ws_bytes and ws.device == occ.device
    with torch.cuda.device(occ.device):
        native.check(L.ffmp_op_flow_field(dev, n, G, C.c_void_p(occ.data_ptr()), C.c_void_p(goals.data_ptr()),
                                          C.c_void_p(cost.data_ptr()) if need_cost else None,
                                          C.c_void_p(flow.data_ptr()), C.c_void_p(ws.data_ptr()), _stream(occ.device)),
                     "ffmp_op_flow_field")
    return (cost if want_cost else None), flow


def flow_field_workspace(n, grid, device):
    """A workspace tensor for flow_field(..., workspace=) on n grids of side `grid`."""
    size = native.lib().ffmp_op_flow_field_workspace(n, grid)
    return torch.empty((max(size, 16),), dtype=torch.uint8, device=device)


def scan(grid_map, pose, beams=360, range_max=3.5, flow_mode=True):
    """LiDAR scan synthesis (SPEC.md §9): grid_map u8[n,G,G] (flow image if flow_mode else occupancy plane), pose f32[n,3]
    -> (ranges f32[n,beams] in metres, +inf = no return, 0 = robot cell occupied; hit u8[n] = FFMP.is_collision2)."""
    L = native.lib()
    dev = _dev_index(grid_map)
    assert grid_map.dtype == torch.uint8 and grid_map.dim() == 3 and grid_map.shape[1] == grid_map.shape[2]
    m = grid_map.contiguous()
    n, G = m.shape[0], m.shape[1]
    p = pose.to(device=m.device, dtype=torch.float32).contiguous()
    assert p.shape == (n, 3)
    out = torch.empty((n, beams), dtype=torch.float32, device=m.device)
    hit = torch.empty((n,), dtype=torch.uint8, device=m.device)
    if n:
        with torch.cuda.device(m.device):
            native.check(L.ffmp_op_scan(dev, n, G, C.c_void_p(m.data_ptr()), 1 if flow_mode else 0, C.c_void_p(p.data_ptr()),
                                        int(beams), C.c_float(range_max), C.c_void_p(out.data_ptr()),
                                        C.c_void_p(hit.data_ptr()), _stream(m.device)), "ffmp_op_scan")
    return out, hit


def flow_dir(flow):
    """Decode direction codes (0..7, 8 = none) from a flow image (SPEC.md §5)."""
    return torch.where(flow == 255, torch.full_like(flow, 8), flow // 28)


def _reward_common(n, device, rel_goal, is_first, d_first):
    rg = rel_goal.to(device=device, dtype=torch.float32).contiguous()
    first = is_first.to(device=device, dtype=torch.uint8).contiguous()
    assert d_first.dtype == torch.float32 and d_first.is_contiguous() and d_first.device == device
    reward = torch.empty((n,), dtype=torch.float32, device=device)
    done = torch.empty((n,), dtype=torch.uint8, device=device)
    flags = torch.empty((n,), dtype=torch.uint8, device=device)
    return rg, first, reward, done, flags


def rewarder(local_map, rel_goal, is_first, d_first):
    """FFMP.rewarder, batched: local_map int32[n,W,W] (>0 occupied, robot at (W/2,W/2)), rel_goal f32[n,2],
    is_first u8[n], d_first f32[n] (latched in place where is_first) -> (reward f32[n], done bool[n], flags u8[n])."""
    L = native.lib()
    dev = _dev_index(local_map)
    lm = local_map.to(torch.int32).contiguous()
    n, W = lm.shape[0], lm.shape[1]
    rg, first, reward, done, flags = _reward_common(n, lm.device, rel_goal, is_first, d_first)
    with torch.cuda.device(lm.device):
        native.check(L.ffmp_op_rewarder(dev, n, W, C.c_void_p(lm.data_ptr()), C.c_void_p(rg.data_ptr()),
                                        C.c_void_p(first.data_ptr()), C.c_void_p(d_first.data_ptr()),
                                        C.c_void_p(reward.data_ptr()), C.c_void_p(done.data_ptr()),
                                        C.c_void_p(flags.data_ptr()), _stream(lm.device)), "ffmp_op_rewarder")
    return reward, done.view(torch.bool), flags


def rewarder2(scan, rel_goal, is_first, d_first):
    """FFMP.rewarder2, batched: scan f64[n,B] LiDAR ranges (NaN = None) instead of a local map."""
    L = native.lib()
    dev = _dev_index(scan)
    sc = scan.to(torch.float64).contiguous()
    n, B = sc.shape
    rg, first, reward, done, flags = _reward_common(n, sc.device, rel_goal, is_first, d_first)
    with torch.cuda.device(sc.device):
        native.check(L.ffmp_op_rewarder2(dev, n, B, C.c_void_p(sc.data_ptr()), C.c_void_p(rg.data_ptr()),
                                         C.c_void_p(first.data_ptr()), C.c_void_p(d_first.data_ptr()),
                                         C.c_void_p(reward.data_ptr()), C.c_void_p(done.data_ptr()),
                                         C.c_void_p(flags.data_ptr()), _stream(sc.device)), "ffmp_op_rewarder2")
    return reward, done.view(torch.bool), flags


def reward_calculator(rel_goal, is_collision, is_goal, is_first, d_first):
    """FFMP.reward_calculator, batched, with explicit collision / goal flags (bool tensors [n])."""
    L = native.lib()
    dev = _dev_index(rel_goal)
    device = rel_goal.device
    n = rel_goal.shape[0]
    given = (is_collision.to(device=device, dtype=torch.uint8) | (is_goal.to(device=device, dtype=torch.uint8) << 1)).contiguous()
    rg, first, reward, done, flags = _reward_common(n, device, rel_goal, is_first, d_first)
    with torch.cuda.device(device):
        native.check(L.ffmp_op_reward_calculator(dev, n, C.c_void_p(rg.data_ptr()), C.c_void_p(given.data_ptr()),
                                                 C.c_void_p(first.data_ptr()), C.c_void_p(d_first.data_ptr()),
                                                 C.c_void_p(reward.data_ptr()), C.c_void_p(done.data_ptr()),
                                                 C.c_void_p(flags.data_ptr()), _stream(device)), "ffmp_op_reward_calculator")
    return reward, done.view(torch.bool), flags
