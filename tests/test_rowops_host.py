"""CPU: the per-row bit-plane arithmetic of the interleaved flow-field kernel (csrc/flow_rowops.cuh), compiled for the host
and driven by a sequential restatement of the kernel's glue (tests/host/rowops_host.cpp), against the oracle.  Covers what the
GPU kernel relies on and nvcc cannot check here: the 32x32 / 16x16 bit-matrix transposes, the PRMT table look-up of the flow
bytes, the one-PRMT-per-cell int32 widening, the Gray-code-by-addition level record and the checkerboard rule for cost bit 0."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import oracle
from maps import special_cases, noise

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "host", "rowops_host.cpp")
LIB = os.path.join(HERE, "host", "librowops_host.so")
HDR = os.path.join(HERE, "..", "flow_field_based_motion_planner_b200", "csrc", "flow_rowops.cuh")


@pytest.fixture(scope="module")
def host():
    if not os.path.exists(LIB) or os.path.getmtime(LIB) < max(os.path.getmtime(SRC), os.path.getmtime(HDR)):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-x", "c++", "-fPIC", "-shared", "-Wall", "-o", LIB, SRC])
    L = C.CDLL(LIB)
    u32p, u8p, i32p = C.POINTER(C.c_uint32), C.POINTER(C.c_uint8), C.POINTER(C.c_int32)
    L.host_transpose32.argtypes = [u32p, u32p]
    L.host_transpose16x2.argtypes = [u32p, u32p]
    L.host_prmt.restype = C.c_uint32
    L.host_prmt.argtypes = [C.c_uint32] * 3
    L.host_occupied_nibbles.restype = C.c_uint32
    L.host_occupied_nibbles.argtypes = [C.c_uint32] * 4
    L.il_emulate_grid.restype = C.c_int
    L.il_emulate_grid.argtypes = [u8p, C.c_int, C.c_int, C.c_int, i32p, u8p]
    L.ilw_emulate_grid.restype = C.c_int
    L.ilw_emulate_grid.argtypes = [u8p, C.c_int, C.c_int, C.c_int, i32p, u8p]
    return L


def _p(a, ty):
    return a.ctypes.data_as(C.POINTER(ty))


def test_transposes(host):
    rng = np.random.default_rng(0)
    x = rng.integers(0, 2**32, 32, dtype=np.uint64).astype(np.uint32)
    out = np.zeros(32, np.uint32)
    host.host_transpose32(_p(x, C.c_uint32), _p(out, C.c_uint32))
    bits = (x[:, None] >> np.arange(32, dtype=np.uint32)[None, :]) & 1          # bits[m, b]
    want = (bits.T.astype(np.uint64) << np.arange(32, dtype=np.uint64)[None, :]).sum(1).astype(np.uint32)
    assert np.array_equal(out, want)
    y = rng.integers(0, 2**32, 16, dtype=np.uint64).astype(np.uint32)
    out16 = np.zeros(16, np.uint32)
    host.host_transpose16x2(_p(y, C.c_uint32), _p(out16, C.c_uint32))
    b16 = (y[:, None] >> np.arange(32, dtype=np.uint32)[None, :]) & 1           # b16[q, bit]
    for m in range(16):
        lo = sum(int(b16[q, m]) << q for q in range(16))
        hi = sum(int(b16[q, 16 + m]) << q for q in range(16))
        assert int(out16[m]) == lo | (hi << 16)


def test_prmt_and_nibbles(host):
    assert host.host_prmt(0x33221100, 0x77665544, 0x7531) == 0x77553311
    assert host.host_prmt(0x80017F00, 0, 0xBA98) == 0xFF000000
    rng = np.random.default_rng(1)
    for _ in range(200):
        by = rng.integers(0, 256, 16, dtype=np.uint64).astype(np.uint8) * (rng.random(16) < 0.5)
        by = by.astype(np.uint8)
        x = by.view(np.uint32)
        z = host.host_occupied_nibbles(int(x[0]), int(x[1]), int(x[2]), int(x[3]))
        for w in range(4):
            for i in range(4):
                assert ((z >> (8 * w + i)) & 1) == int(by[4 * i + w] != 0)
        assert z & 0xF0F0F0F0 == 0


def _check(host, occ, goal):
    G = occ.shape[0]
    occ = np.ascontiguousarray(occ, np.uint8)
    cost = np.zeros((G, G), np.int32)
    flow = np.zeros((G, G), np.uint8)
    host.il_emulate_grid(_p(occ, C.c_uint8), G, int(goal[0]), int(goal[1]), _p(cost, C.c_int32), _p(flow, C.c_uint8))
    ec, _, ef = oracle.flow_field(occ, goal[0], goal[1])
    assert np.array_equal(cost, ec)
    assert np.array_equal(flow, ef)


@pytest.mark.parametrize("G", [100, 112, 124, 128])
def test_interleaved_algorithm_matches_oracle_on_special_maps(host, G):
    for name, occ, goal in special_cases(G):
        _check(host, occ, goal)


def test_interleaved_algorithm_matches_oracle_on_generated_maps(host):
    for env in range(24):
        occ, _, _, cells = oracle.scenario(11, env, env % 5, 128, p_occ=0.1 + 0.02 * (env % 4), block_shift=env % 4)
        _check(host, occ, (cells[2], cells[3]))
    for sd in range(6):
        occ, g = noise(128, 0.3, sd, values=(1, 255, 128))
        _check(host, occ, g)


def _check_wide(host, occ, goal):
    G = occ.shape[0]
    occ = np.ascontiguousarray(occ, np.uint8)
    cost = np.zeros((G, G), np.int32)
    flow = np.zeros((G, G), np.uint8)
    host.ilw_emulate_grid(_p(occ, C.c_uint8), G, int(goal[0]), int(goal[1]), _p(cost, C.c_int32), _p(flow, C.c_uint8))
    ec, _, ef = oracle.flow_field(occ, goal[0], goal[1])
    assert np.array_equal(flow, ef)
    assert np.array_equal(cost, ec)


@pytest.mark.parametrize("G", [416, 512])
def test_wide_row_algorithm_matches_oracle(host, G):
    """16 interleaved words per row (flow_field_wide.cu): two-byte cost widening, the 16 x 32 flag transpose of the input."""
    for name, occ, goal in special_cases(G):
        if name in ("serpentine", "rooms1", "rooms2", "noise0.05", "noise0.45", "goal_out_of_grid_hi"):
            continue                  # keep the CPU suite short: one of each kind stays
        _check_wide(host, occ, goal)
    occ, _, _, cells = oracle.scenario(3, 1, 0, G, p_occ=0.3, block_shift=0)
    _check_wide(host, occ, (cells[2], cells[3]))
