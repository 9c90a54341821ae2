// flow_field.cu — SPEC.md §4/§5: integration field (F1), flow direction (F2) and flow image, G <= 128.
// Replaces the external /bev/* flow-image ROS node (/root/reference/src/train.py:84,116-121).
//
// Design (one WARP per grid, no block barrier anywhere):
//   * the occupancy plane is staged into shared memory with one TMA bulk copy (cp.async.bulk +
//     mbarrier) and packed to a bit mask: lane l owns rows [l*RPL, l*RPL+RPL) as RPL x WPR 32-bit
//     words held in REGISTERS (G=128: 4 rows x 4 words);
//   * the wavefront is bit-parallel: new = (west|east|north|south of frontier) & avail; left/right
//     neighbours are funnel shifts, up/down neighbours are the adjacent row registers (one warp
//     shuffle per boundary row); convergence is a warp vote (__any_sync);
//   * the BFS level of a cell is recorded as Gray-code BIT-PLANES: Gray(L) differs from Gray(L-1)
//     in bit ctz(L) only, so level L costs one XOR of `avail` into plane ctz(L) — no per-cell work
//     inside the level loop.  Planes 0..7 live in shared memory (levels < 256), higher planes spill
//     to a per-CTA global scratch (rare: mazes);
//   * afterwards the planes are un-Gray'd in place; the 8-neighbour argmin is evaluated bit-parallel
//     from cost bits 0..2 (adjacent free cells differ by exactly +-1, admissible diagonals by 0/+-2),
//     reproducing the scan order E,NE,N,NW,W,SW,S,SE with strict '<';
//   * cost (int32) and the flow image (u8) are expanded from the bit-planes and stored.
// Algorithmic HBM bytes: 6 B/cell (1 occ read + 4 cost write + 1 flow write).
#include "ffmp_kernels.cuh"

namespace ffmp {

namespace {

constexpr int NPL = 8;  // bit-planes resident in shared memory

template <int WPR>
struct Row {
    static __device__ __forceinline__ void ld(const uint32_t *p, uint32_t (&v)[WPR]) {
        if constexpr (WPR == 4) {
            const uint4 t = *reinterpret_cast<const uint4 *>(p);
            v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
        } else if constexpr (WPR == 2) {
            const uint2 t = *reinterpret_cast<const uint2 *>(p);
            v[0] = t.x; v[1] = t.y;
        } else {
#pragma unroll
            for (int w = 0; w < WPR; ++w) v[w] = p[w];
        }
    }
    static __device__ __forceinline__ void st(uint32_t *p, const uint32_t (&v)[WPR]) {
        if constexpr (WPR == 4) {
            *reinterpret_cast<uint4 *>(p) = make_uint4(v[0], v[1], v[2], v[3]);
        } else if constexpr (WPR == 2) {
            *reinterpret_cast<uint2 *>(p) = make_uint2(v[0], v[1]);
        } else {
#pragma unroll
            for (int w = 0; w < WPR; ++w) p[w] = v[w];
        }
    }
};

// value of the cell at column+1 / column-1 aligned to this word
template <int WPR>
__device__ __forceinline__ uint32_t shr1(const uint32_t (&x)[WPR], int w) {
    return (x[w] >> 1) | (w + 1 < WPR ? x[w + 1] << 31 : 0u);
}
template <int WPR>
__device__ __forceinline__ uint32_t shl1(const uint32_t (&x)[WPR], int w) {
    return (x[w] << 1) | (w > 0 ? x[w - 1] >> 31 : 0u);
}

// spread the 4 bits of nibble n of x to the low bit of 4 bytes
__device__ __forceinline__ uint32_t spread4(uint32_t x, int n) {
    return (((x >> (4 * n)) & 0xFu) * 0x00204081u) & 0x01010101u;
}

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// 8x8 bit-matrix transpose of the 64-bit value hi:lo (Hacker's Delight 7-3), on 32-bit halves
__device__ __forceinline__ void transpose8(uint32_t &lo, uint32_t &hi) {
    uint32_t t;
    t = (lo ^ (lo >> 7)) & 0x00AA00AAu; lo ^= t ^ (t << 7);
    t = (hi ^ (hi >> 7)) & 0x00AA00AAu; hi ^= t ^ (t << 7);
    t = (lo ^ (lo >> 14)) & 0x0000CCCCu; lo ^= t ^ (t << 14);
    t = (hi ^ (hi >> 14)) & 0x0000CCCCu; hi ^= t ^ (t << 14);
    t = (lo ^ __funnelshift_r(lo, hi, 28)) & 0xF0F0F0F0u;
    lo ^= t; hi ^= t >> 4;
}

// 4x4 byte transpose: out[b] = (p0.b, p1.b, p2.b, p3.b)
__device__ __forceinline__ void bytes4x4(uint32_t p0, uint32_t p1, uint32_t p2, uint32_t p3, uint32_t (&out)[4]) {
    const uint32_t a = __byte_perm(p0, p1, 0x5140), b = __byte_perm(p0, p1, 0x7362);
    const uint32_t c = __byte_perm(p2, p3, 0x5140), d = __byte_perm(p2, p3, 0x7362);
    out[0] = __byte_perm(a, c, 0x5410); out[1] = __byte_perm(a, c, 0x7632);
    out[2] = __byte_perm(b, d, 0x5410); out[3] = __byte_perm(b, d, 0x7632);
}

// GEN = true : the scenario (SPEC.md §3) is generated in-kernel from the hash RNG straight into the bit
//              mask (no occupancy plane round trip); used by the batched env (reset and regeneration).
// GEN = false: the occupancy plane is an input (stateless operator), staged with one TMA bulk copy.
template <int WPR, bool GEN>
__global__ void __launch_bounds__(32) flow_field_warp_kernel(FlowArgs a) {
    constexpr int RPL = WPR;                    // rows per lane
    constexpr int PLANE_WORDS = 32 * RPL * WPR;  // one bit-plane of the padded (32*WPR)^2 grid
    __shared__ __align__(128) uint32_t pl[NPL * PLANE_WORDS];
    __shared__ __align__(8) uint64_t mbar;

    const int lane = threadIdx.x;
    const int G = a.G;
    const int count = a.count_ptr ? static_cast<int>(*a.count_ptr) : a.count;
    const uint32_t bar = static_cast<uint32_t>(__cvta_generic_to_shared(&mbar));
    const uint32_t pl_s = static_cast<uint32_t>(__cvta_generic_to_shared(pl));
    uint32_t *hi = a.hi_scratch + static_cast<size_t>(blockIdx.x) * (8 * PLANE_WORDS);
    uint32_t parity = 0;

    if (!GEN) {
        if (lane == 0) {
            mbar_init(bar, 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncwarp();
    }

    auto pidx = [&](int k, int r, int ln) { return ((k * RPL + r) * 32 + ln) * WPR; };

    for (int item = blockIdx.x; item < count; item += gridDim.x) {
        const uint32_t env = a.env_idx ? a.env_idx[item] : static_cast<uint32_t>(item);
        const size_t cells = static_cast<size_t>(G) * G;
        size_t plane;
        int gi, gj;
        uint32_t FR[RPL][WPR];

        if (GEN) {
            // ---- 1g. scenario parameters (lane 0) and the free-cell mask straight from the hash ---
            const uint32_t episode = a.episode ? a.episode[item] : a.episode_const;
            plane = static_cast<size_t>(episode % a.S) * a.N + env;
            const uint32_t key = scenario_key(a.seed, a.env_id_base + env, episode);
            ScenarioParams sp;
            sp.si = sp.sj = sp.gi = sp.gj = 0; sp.yaw = 0.0f;
            if (lane == 0) {
                sp = sample_scenario(key, G, a.goal_mode);
                store_scenario_record(a.scen_out + plane * SC_WORDS, sp, key);
            }
            sp.si = __shfl_sync(FULL, sp.si, 0); sp.sj = __shfl_sync(FULL, sp.sj, 0);
            sp.gi = __shfl_sync(FULL, sp.gi, 0); sp.gj = __shfl_sync(FULL, sp.gj, 0);
            gi = sp.gi; gj = sp.gj;
#pragma unroll
            for (int r = 0; r < RPL; ++r)
#pragma unroll
                for (int w = 0; w < WPR; ++w)
                    FR[r][w] = scenario_free_word(key, lane * RPL + r, 32 * w, G, a.block_shift, a.p_thresh, sp);
        } else {
            if (a.slot_mode) {
                plane = static_cast<size_t>((a.episode ? a.episode[item] : a.episode_const) % a.S) * a.N + env;
                gi = static_cast<int>(a.scen[plane * SC_WORDS + SC_GI]);
                gj = static_cast<int>(a.scen[plane * SC_WORDS + SC_GJ]);
            } else {
                plane = static_cast<size_t>(item);
                gi = a.goal_cells[2 * item];
                gj = a.goal_cells[2 * item + 1];
            }
            // ---- 1. TMA bulk copy of the occupancy plane into shared memory ----------------------
            fence_proxy_async();  // earlier generic-proxy accesses to `pl` are ordered before the async write
            __syncwarp();
            if (lane == 0) {
                mbar_expect_tx(bar, static_cast<uint32_t>(cells));
                tma_bulk_g2s(pl_s, a.occ + plane * cells, static_cast<uint32_t>(cells), bar);
            }
            mbar_wait(bar, parity);
            parity ^= 1;

            // ---- 2. bytes -> free-cell bit mask (registers) --------------------------------------
            if (G == 32 * WPR) {
                // bit packing is linear in the flat cell index: every lane packs 16 consecutive bytes into
                // 16 bits (conflict-free LDS.128), written in place at byte f/8; each lane then owns the
                // WPR*WPR consecutive words of its rows
                uint8_t *stage = reinterpret_cast<uint8_t *>(pl);
#pragma unroll 1
                for (int it = 0; it < (32 * WPR * WPR) / 16; ++it) {
                    const int f = 16 * (32 * it + lane);
                    const uint4 q = *reinterpret_cast<const uint4 *>(stage + f);
                    const uint32_t x4[4] = {q.x, q.y, q.z, q.w};
                    uint32_t bits = 0;
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const uint32_t x = x4[u];
                        const uint32_t nz = (x | ((x & 0x7F7F7F7Fu) + 0x7F7F7F7Fu)) & 0x80808080u;   // 0x80 per non-zero byte
                        const uint32_t fr = (~nz >> 7) & 0x01010101u;
                        bits |= ((fr * 0x01020408u) >> 24 & 0xFu) << (4 * u);
                    }
                    __syncwarp();
                    *reinterpret_cast<uint16_t *>(stage + (f >> 3)) = static_cast<uint16_t>(bits);
                    __syncwarp();
                }
#pragma unroll
                for (int r = 0; r < RPL; ++r) Row<WPR>::ld(&pl[(lane * RPL + r) * WPR], FR[r]);
            } else {
                const uint8_t *stage = reinterpret_cast<const uint8_t *>(pl);
#pragma unroll 1
                for (int o = 0; o < 32; ++o) {
#pragma unroll
                    for (int r = 0; r < RPL; ++r) {
                        const int R = o * RPL + r;
#pragma unroll
                        for (int w = 0; w < WPR; ++w) {
                            const int col = 32 * w + lane;
                            const bool fr = (R < G && col < G) ? stage[R * G + col] == 0 : false;
                            const uint32_t bits = __ballot_sync(FULL, fr);
                            if (lane == o) FR[r][w] = bits;
                        }
                    }
                }
            }
            __syncwarp();
        }

        // ---- 3. zero the resident bit-planes 2.. (planes 0 and 1 live in registers during the BFS) ----
        {
            uint32_t z[WPR];
#pragma unroll
            for (int w = 0; w < WPR; ++w) z[w] = 0;
#pragma unroll
            for (int k = 2; k < NPL; ++k)
#pragma unroll
                for (int r = 0; r < RPL; ++r) Row<WPR>::st(&pl[pidx(k, r, lane)], z);
        }

        // ---- 4. bit-parallel wavefront -----------------------------------------------------------
        uint32_t A[RPL][WPR], F[RPL][WPR], G0[RPL][WPR], G1[RPL][WPR];
#pragma unroll
        for (int r = 0; r < RPL; ++r)
#pragma unroll
            for (int w = 0; w < WPR; ++w) { A[r][w] = FR[r][w]; F[r][w] = 0; G0[r][w] = 0; G1[r][w] = 0; }
        if (gi >= 0 && gj >= 0 && gi < G && gj < G && lane == gi / RPL) {
#pragma unroll
            for (int r = 0; r < RPL; ++r)
#pragma unroll
                for (int w = 0; w < WPR; ++w)
                    if (r == gi % RPL && w == (gj >> 5)) {
                        const uint32_t m = (1u << (gj & 31)) & A[r][w];
                        F[r][w] = m;
                        A[r][w] &= ~m;
                    }
        }
        const uint32_t upmask = lane == 0 ? 0u : 0xFFFFFFFFu, dnmask = lane == 31 ? 0u : 0xFFFFFFFFu;
        uint32_t L = 1;
        for (;; ++L) {
            // Gray bit-plane update: cells with cost >= L flip Gray bit ctz(L)
            const int k = __ffs(L) - 1;
            if (k == 0) {
#pragma unroll
                for (int r = 0; r < RPL; ++r)
#pragma unroll
                    for (int w = 0; w < WPR; ++w) G0[r][w] ^= A[r][w];
            } else if (k == 1) {
#pragma unroll
                for (int r = 0; r < RPL; ++r)
#pragma unroll
                    for (int w = 0; w < WPR; ++w) G1[r][w] ^= A[r][w];
            } else if (k < NPL) {
#pragma unroll
                for (int r = 0; r < RPL; ++r) {
                    uint32_t v[WPR];
                    uint32_t *p = &pl[pidx(k, r, lane)];
                    Row<WPR>::ld(p, v);
#pragma unroll
                    for (int w = 0; w < WPR; ++w) v[w] ^= A[r][w];
                    Row<WPR>::st(p, v);
                }
            } else {
                const bool first = L == (1u << k);
#pragma unroll
                for (int r = 0; r < RPL; ++r)
#pragma unroll
                    for (int w = 0; w < WPR; ++w) {
                        uint32_t *p = &hi[pidx(k - NPL, r, lane) + w];
                        *p = first ? A[r][w] : (*p ^ A[r][w]);
                    }
            }
            uint32_t upF[WPR], dnF[WPR];
#pragma unroll
            for (int w = 0; w < WPR; ++w) {
                upF[w] = __shfl_up_sync(FULL, F[RPL - 1][w], 1) & upmask;
                dnF[w] = __shfl_down_sync(FULL, F[0][w], 1) & dnmask;
            }
            uint32_t Nw[RPL][WPR];
            uint32_t any = 0;
#pragma unroll
            for (int r = 0; r < RPL; ++r)
#pragma unroll
                for (int w = 0; w < WPR; ++w) {
                    const uint32_t up = r == 0 ? upF[w] : F[r - 1][w];
                    const uint32_t dn = r == RPL - 1 ? dnF[w] : F[r + 1][w];
                    const uint32_t n = (shl1<WPR>(F[r], w) | shr1<WPR>(F[r], w) | up | dn) & A[r][w];
                    Nw[r][w] = n;
                    any |= n;
                }
#pragma unroll
            for (int r = 0; r < RPL; ++r)
#pragma unroll
                for (int w = 0; w < WPR; ++w) { A[r][w] &= ~Nw[r][w]; F[r][w] = Nw[r][w]; }
            if (!__any_sync(FULL, any != 0)) break;
        }
#pragma unroll
        for (int r = 0; r < RPL; ++r) {
            Row<WPR>::st(&pl[pidx(0, r, lane)], G0[r]);
            Row<WPR>::st(&pl[pidx(1, r, lane)], G1[r]);
        }
        const uint32_t Lmax = L - 1;                       // deepest level that reached a cell
        const int kmax = 32 - __clz(Lmax);                  // number of significant cost bits
        // visited (= reached free) and free masks go to scratch planes 14/15 so that the remaining
        // phases can be ROLLED loops over (row, word) reading memory: keeps the code inside the I-cache
        constexpr int PVIS = 14, PFREE = 15;
#pragma unroll
        for (int r = 0; r < RPL; ++r) {
            uint32_t v[WPR];
#pragma unroll
            for (int w = 0; w < WPR; ++w) v[w] = FR[r][w] & ~A[r][w];
            Row<WPR>::st(&hi[pidx(PVIS - NPL, r, lane)], v);
            Row<WPR>::st(&hi[pidx(PFREE - NPL, r, lane)], FR[r]);
        }
        __syncwarp();

        // ---- 5. Gray -> binary, in place ---------------------------------------------------------
#pragma unroll 1
        for (int r = 0; r < RPL; ++r) {
            uint32_t acc[WPR];
#pragma unroll
            for (int w = 0; w < WPR; ++w) acc[w] = 0;
#pragma unroll 1
            for (int k = kmax - 1; k >= 0; --k) {
                uint32_t v[WPR];
                uint32_t *p = k < NPL ? &pl[pidx(k, r, lane)] : &hi[pidx(k - NPL, r, lane)];
                Row<WPR>::ld(p, v);
#pragma unroll
                for (int w = 0; w < WPR; ++w) { acc[w] ^= v[w]; v[w] = acc[w]; }
                Row<WPR>::st(p, v);
            }
        }
        __syncwarp();

        // ---- 6. expand the integration field to int32 and store ---------------------------------
        if (a.cost) {
            int32_t *cost = a.cost + plane * cells;
            if (kmax <= NPL) {
                // fast path (depth < 256): 8 planes x 32 cells -> 32 bytes with 4x4 byte transposes (PRMT) and
                // 8x8 bit-matrix transposes, then widened to int32 with INF for unreached cells
#pragma unroll 1
                for (int rw = 0; rw < RPL * WPR; ++rw) {
                    const int r = rw / WPR, w = rw - r * WPR;
                    const int R = lane * RPL + r;
                    if (R >= G || 32 * w >= G) continue;
                    uint32_t P[NPL];
#pragma unroll
                    for (int k = 0; k < NPL; ++k) P[k] = k < kmax ? pl[pidx(k, r, lane) + w] : 0u;
                    const uint32_t vis = hi[pidx(PVIS - NPL, r, lane) + w];
                    uint32_t tl[4], th[4];
                    bytes4x4(P[0], P[1], P[2], P[3], tl);
                    bytes4x4(P[4], P[5], P[6], P[7], th);
                    int32_t *dst = cost + static_cast<size_t>(R) * G + 32 * w;
#pragma unroll
                    for (int b = 0; b < 4; ++b) {
                        if (32 * w + 8 * b >= G) continue;
                        uint32_t lo = tl[b], hb = th[b];
                        transpose8(lo, hb);          // byte j of hb:lo = cost of cell 8b+j
                        int4 c0, c1;
                        c0.x = (vis >> (8 * b + 0)) & 1 ? static_cast<int>(lo & 0xFF) : COST_INF;
                        c0.y = (vis >> (8 * b + 1)) & 1 ? static_cast<int>((lo >> 8) & 0xFF) : COST_INF;
                        c0.z = (vis >> (8 * b + 2)) & 1 ? static_cast<int>((lo >> 16) & 0xFF) : COST_INF;
                        c0.w = (vis >> (8 * b + 3)) & 1 ? static_cast<int>(lo >> 24) : COST_INF;
                        c1.x = (vis >> (8 * b + 4)) & 1 ? static_cast<int>(hb & 0xFF) : COST_INF;
                        c1.y = (vis >> (8 * b + 5)) & 1 ? static_cast<int>((hb >> 8) & 0xFF) : COST_INF;
                        c1.z = (vis >> (8 * b + 6)) & 1 ? static_cast<int>((hb >> 16) & 0xFF) : COST_INF;
                        c1.w = (vis >> (8 * b + 7)) & 1 ? static_cast<int>(hb >> 24) : COST_INF;
                        *reinterpret_cast<int4 *>(dst + 8 * b) = c0;
                        if (32 * w + 8 * b + 4 < G) *reinterpret_cast<int4 *>(dst + 8 * b + 4) = c1;
                    }
                }
            } else {
#pragma unroll 1
                for (int rw = 0; rw < RPL * WPR; ++rw) {
                    const int r = rw / WPR, w = rw - r * WPR;
                    const int R = lane * RPL + r;
                    if (R >= G || 32 * w >= G) continue;
                    uint32_t lo[8], hb[8];
#pragma unroll
                    for (int n = 0; n < 8; ++n) { lo[n] = 0; hb[n] = 0; }
#pragma unroll 1
                    for (int k = 0; k < NPL; ++k) {
                        const uint32_t word = pl[pidx(k, r, lane) + w];
#pragma unroll
                        for (int n = 0; n < 8; ++n) lo[n] += spread4(word, n) << k;
                    }
#pragma unroll 1
                    for (int k = NPL; k < kmax; ++k) {
                        const uint32_t word = hi[pidx(k - NPL, r, lane) + w];
#pragma unroll
                        for (int n = 0; n < 8; ++n) hb[n] += spread4(word, n) << (k - NPL);
                    }
                    const uint32_t vis = hi[pidx(PVIS - NPL, r, lane) + w];
                    int32_t *dst = cost + static_cast<size_t>(R) * G + 32 * w;
#pragma unroll
                    for (int n = 0; n < 8; ++n) {
                        if (32 * w + 4 * n >= G) continue;
                        int4 c;
                        c.x = (vis >> (4 * n + 0)) & 1 ? static_cast<int>((lo[n] & 0xFF) | ((hb[n] & 0xFF) << 8)) : COST_INF;
                        c.y = (vis >> (4 * n + 1)) & 1 ? static_cast<int>(((lo[n] >> 8) & 0xFF) | (((hb[n] >> 8) & 0xFF) << 8)) : COST_INF;
                        c.z = (vis >> (4 * n + 2)) & 1 ? static_cast<int>(((lo[n] >> 16) & 0xFF) | (((hb[n] >> 16) & 0xFF) << 8)) : COST_INF;
                        c.w = (vis >> (4 * n + 3)) & 1 ? static_cast<int>((lo[n] >> 24) | ((hb[n] >> 24) << 8)) : COST_INF;
                        *reinterpret_cast<int4 *>(dst + 4 * n) = c;
                    }
                }
            }
        }
        __syncwarp();

        // ---- 7. flow direction, bit-parallel -----------------------------------------------------
        // Resident planes 3..7 are free now: 3 = visited, 4 = free (neighbour rows are read back from
        // other lanes), 5/6/7 + plane 0 (own row only) receive the 4 direction-code bit-planes.
        constexpr int PV = 3, PF = 4;
#pragma unroll 1
        for (int r = 0; r < RPL; ++r) {
            uint32_t v[WPR];
            Row<WPR>::ld(&hi[pidx(PVIS - NPL, r, lane)], v);
            Row<WPR>::st(&pl[pidx(PV, r, lane)], v);
            Row<WPR>::ld(&hi[pidx(PFREE - NPL, r, lane)], v);
            Row<WPR>::st(&pl[pidx(PF, r, lane)], v);
        }
        __syncwarp();
#pragma unroll 1
        for (int r = 0; r < RPL; ++r) {
            // rows R-1 (west, "u") and R+1 (east, "d")
            const int lu = r == 0 ? lane - 1 : lane, ru = r == 0 ? RPL - 1 : r - 1;
            const int ld = r == RPL - 1 ? lane + 1 : lane, rd = r == RPL - 1 ? 0 : r + 1;
            uint32_t b0[WPR], b1c[WPR], b2c[WPR], b1u[WPR], b2u[WPR], b1d[WPR], b2d[WPR];
            uint32_t Vc[WPR], Vu[WPR], Vd[WPR], Fc[WPR], Fu[WPR], Fd[WPR];
#pragma unroll
            for (int w = 0; w < WPR; ++w) {
                b0[w] = b1c[w] = b2c[w] = b1u[w] = b2u[w] = b1d[w] = b2d[w] = 0;
                Vu[w] = Vd[w] = Fu[w] = Fd[w] = 0;
            }
            if (kmax > 0) Row<WPR>::ld(&pl[pidx(0, r, lane)], b0);
            if (kmax > 1) Row<WPR>::ld(&pl[pidx(1, r, lane)], b1c);
            if (kmax > 2) Row<WPR>::ld(&pl[pidx(2, r, lane)], b2c);
            Row<WPR>::ld(&pl[pidx(PV, r, lane)], Vc);
            Row<WPR>::ld(&pl[pidx(PF, r, lane)], Fc);
            if (lu >= 0) {
                if (kmax > 1) Row<WPR>::ld(&pl[pidx(1, ru, lu)], b1u);
                if (kmax > 2) Row<WPR>::ld(&pl[pidx(2, ru, lu)], b2u);
                Row<WPR>::ld(&pl[pidx(PV, ru, lu)], Vu);
                Row<WPR>::ld(&pl[pidx(PF, ru, lu)], Fu);
            }
            if (ld < 32) {
                if (kmax > 1) Row<WPR>::ld(&pl[pidx(1, rd, ld)], b1d);
                if (kmax > 2) Row<WPR>::ld(&pl[pidx(2, rd, ld)], b2d);
                Row<WPR>::ld(&pl[pidx(PV, rd, ld)], Vd);
                Row<WPR>::ld(&pl[pidx(PF, rd, ld)], Fd);
            }
            uint32_t d0[WPR], d1[WPR], d2[WPR], d3[WPR];
#pragma unroll
            for (int w = 0; w < WPR; ++w) {
                const uint32_t own = Vc[w];
                const uint32_t t = b1c[w] ^ ~b0[w];   // bit 1 of (cost-1)
                const uint32_t u = b2c[w] ^ ~b1c[w];  // bit 2 of (cost-2)
                // orthogonal neighbours one level lower (codes 0 E, 2 N, 4 W, 6 S)
                const uint32_t lE = own & Vd[w] & ~(b1d[w] ^ t);
                const uint32_t lW = own & Vu[w] & ~(b1u[w] ^ t);
                const uint32_t lN = own & shr1<WPR>(Vc, w) & ~(shr1<WPR>(b1c, w) ^ t);
                const uint32_t lS = own & shl1<WPR>(Vc, w) & ~(shl1<WPR>(b1c, w) ^ t);
                // admissible diagonals two levels lower (codes 1 NE, 3 NW, 5 SW, 7 SE)
                const uint32_t fE = Fd[w], fW = Fu[w], fN = shr1<WPR>(Fc, w), fS = shl1<WPR>(Fc, w);
                const uint32_t lNE = own & shr1<WPR>(Fd, w) & fE & fN & (shr1<WPR>(b1d, w) ^ b1c[w]) & ~(shr1<WPR>(b2d, w) ^ u);
                const uint32_t lNW = own & shr1<WPR>(Fu, w) & fW & fN & (shr1<WPR>(b1u, w) ^ b1c[w]) & ~(shr1<WPR>(b2u, w) ^ u);
                const uint32_t lSW = own & shl1<WPR>(Fu, w) & fW & fS & (shl1<WPR>(b1u, w) ^ b1c[w]) & ~(shl1<WPR>(b2u, w) ^ u);
                const uint32_t lSE = own & shl1<WPR>(Fd, w) & fE & fS & (shl1<WPR>(b1d, w) ^ b1c[w]) & ~(shl1<WPR>(b2d, w) ^ u);
                const uint32_t anyD = lNE | lNW | lSW | lSE;
                const uint32_t m0 = (anyD & lNE) | (~anyD & lE);
                const uint32_t m1 = (anyD & lNW) | (~anyD & lN);
                const uint32_t m2 = (anyD & lSW) | (~anyD & lW);
                const uint32_t m3 = (anyD & lSE) | (~anyD & lS);
                d0[w] = anyD;
                d1[w] = ~m0 & (m1 | (~m2 & m3));
                d2[w] = ~m0 & ~m1 & (m2 | m3);
                d3[w] = ~(m0 | m1 | m2 | m3);
            }
            // b0 of this row is not read by any other lane: the row can be overwritten right away
            Row<WPR>::st(&pl[pidx(0, r, lane)], d0);
            Row<WPR>::st(&pl[pidx(5, r, lane)], d1);
            Row<WPR>::st(&pl[pidx(6, r, lane)], d2);
            Row<WPR>::st(&pl[pidx(7, r, lane)], d3);
        }
        __syncwarp();

        // ---- 8. expand the flow image (255 occupied, else dir*28) and store ----------------------
        uint8_t *flow = a.flow + plane * cells;
#pragma unroll 1
        for (int rw = 0; rw < RPL * WPR; ++rw) {
            const int r = rw / WPR, w = rw - r * WPR;
            const int R = lane * RPL + r;
            if (R >= G || 32 * w >= G) continue;
            const uint32_t d0 = pl[pidx(0, r, lane) + w], d1 = pl[pidx(5, r, lane) + w];
            const uint32_t d2 = pl[pidx(6, r, lane) + w], d3 = pl[pidx(7, r, lane) + w];
            const uint32_t occ = ~pl[pidx(PF, r, lane) + w];
            uint8_t *dst = flow + static_cast<size_t>(R) * G + 32 * w;
#pragma unroll
            for (int n = 0; n < 8; ++n) {
                if (32 * w + 4 * n >= G) continue;
                uint32_t v = spread4(d0, n) * 28u + spread4(d1, n) * 56u + spread4(d2, n) * 112u + spread4(d3, n) * 224u;
                v |= spread4(occ, n) * 255u;
                *reinterpret_cast<uint32_t *>(dst + 4 * n) = v;
            }
        }
        __syncwarp();
    }

    // the last CTA to finish re-arms the regeneration list for its next use
    if (a.ticket && lane == 0) {
        __threadfence();
        const uint32_t t = atomicAdd(a.ticket, 1u);
        if (t == gridDim.x - 1) {
            *a.ticket = 0;
            if (a.count_reset) *a.count_reset = 0;
            __threadfence();
        }
    }
}

}  // namespace

bool flow_field_large_supported(int G);
int flow_field_large_max_grid(int G);
cudaError_t launch_flow_field_large(const FlowArgs &a, int grid, cudaStream_t st);

bool flow_field_supported(int G) { return (G >= 16 && G <= 128 && (G % 4) == 0) || flow_field_large_supported(G); }

size_t flow_field_scratch_words(int G) {
    if (G > 128) return 0;   // the large-map kernel keeps everything in shared memory and the cost plane
    // spill planes for cost bits 8..15 of the padded grid ((G+31)/32*32)^2
    const int wpr = (G + 31) / 32;
    return static_cast<size_t>(8) * 32 * wpr * wpr;
}

int flow_field_max_grid(int G) {
    if (G > 128) return flow_field_large_max_grid(G);
    const int wpr = (G + 31) / 32;
    const int smem = NPL * 32 * wpr * wpr * 4 + 1024;
    int per_sm = (227 * 1024) / smem;
    if (per_sm > 32) per_sm = 32;
    return 148 * per_sm;
}

cudaError_t launch_flow_field(const FlowArgs &a, int grid, cudaStream_t st) {
    if (grid <= 0) return cudaSuccess;
    if (a.G > 128) return launch_flow_field_large(a, grid, st);
    const int wpr = (a.G + 31) / 32;
    if (a.generate) {
        switch (wpr) {
        case 1: flow_field_warp_kernel<1, true><<<grid, 32, 0, st>>>(a); break;
        case 2: flow_field_warp_kernel<2, true><<<grid, 32, 0, st>>>(a); break;
        case 3: flow_field_warp_kernel<3, true><<<grid, 32, 0, st>>>(a); break;
        case 4: flow_field_warp_kernel<4, true><<<grid, 32, 0, st>>>(a); break;
        default: return cudaErrorInvalidValue;
        }
    } else {
        switch (wpr) {
        case 1: flow_field_warp_kernel<1, false><<<grid, 32, 0, st>>>(a); break;
        case 2: flow_field_warp_kernel<2, false><<<grid, 32, 0, st>>>(a); break;
        case 3: flow_field_warp_kernel<3, false><<<grid, 32, 0, st>>>(a); break;
        case 4: flow_field_warp_kernel<4, false><<<grid, 32, 0, st>>>(a); break;
        default: return cudaErrorInvalidValue;
        }
    }
    return cudaGetLastError();
}

}  // namespace ffmp
