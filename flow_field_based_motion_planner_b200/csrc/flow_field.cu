// flow_field.cu — SPEC.md §4/§5: integration field (F1), flow direction (F2) and flow image, G <= 96 (linear bit layout;
// 96 < G <= 128 is served by flow_field_il.cu, 128 < G <= 512 by flow_field_large.cu) and the launcher of all three.
// Replaces the external /bev/* flow-image ROS node (/root/reference/src/train.py:84,116-121).
//
// Design (one WARP per grid, no block barrier anywhere):
//   * GEN: the scenario (SPEC.md §3) is generated from the hash RNG straight into the bit mask; otherwise the
//     occupancy plane is staged into shared memory with one TMA bulk copy (cp.async.bulk + mbarrier) and packed;
//     lane l owns rows [l*RPL, l*RPL+RPL) as RPL x WPR 32-bit words held in REGISTERS (G=128: 4 rows x 4 words);
//   * the wavefront is bit-parallel: new = (west|east|north|south of frontier) & avail; left/right neighbours are
//     funnel shifts (SHF.L.W / SHF.R.W), up/down neighbours are the adjacent row registers (one warp shuffle per
//     boundary row).  The loop is ALU-pipe bound (LOP3/SHF issue every other cycle per scheduler, measured in
//     profiles/r01a_pipe_probe.txt), so everything that is not a logic op is pushed to the FMA pipe: the
//     avail update A -= new and the boundary-lane masks are IMADs with an opaque multiplier;
//   * the BFS level of a cell is recorded as Gray-code BIT-PLANES: Gray(L) differs from Gray(L-1) in bit ctz(L)
//     only, so level L costs one XOR of `avail` into plane ctz(L) — no per-cell work inside the level loop.  The
//     loop is unrolled by four levels (plane index static for three of them) and tests convergence (warp vote)
//     once per four levels.  Planes 0/1 live in registers, 2..7 in shared memory (levels < 256), higher planes
//     spill to a per-CTA global scratch (rare: mazes);
//   * afterwards the planes are un-Gray'd in place; the 8-neighbour argmin is evaluated bit-parallel from cost
//     bits 0..2 (adjacent free cells differ by exactly +-1, admissible diagonals by 0/+-2), reproducing the scan
//     order E,NE,N,NW,W,SW,S,SE with strict '<'; direction bit-planes become flow bytes with PRMT / 8x8 bit
//     transposes; the cost planes become bytes the same way, are staged (XOR-swizzled) in the plane storage they
//     came from, and are widened to int32 with fully coalesced 16-byte stores.
// Algorithmic HBM bytes: 6 B/cell (1 occ read + 4 cost write + 1 flow write).
#include <cstdlib>
#include <type_traits>

#include "flow_bits.cuh"

namespace ffmp {

namespace {

constexpr int NPL = 8;   // cost bit-planes of the byte fast path (depth < 256)
// Bit-planes 0..NPS-1 are resident in shared memory, planes NPS..15 live in the per-CTA global scratch (L2).  Six resident
// planes (16 KB per warp with the visited / free planes at G = 128): the operator needs P*P bytes of staging for the occupancy
// plane anyway.  The in-kernel-generation variant used to keep 4 (12 KB, a smaller footprint next to the env's step kernel),
// but the step is work-conserving — what counts is footprint x residence time — and with 6 planes a regeneration launch takes
// 87 us instead of 115 us (no L2 round trips every 16th level): same step time at 8 scenario slots, 10 % better at 6.
__host__ __device__ constexpr int nps_of(bool) { return 6; }
constexpr int NPG = 16 - 4;   // scratch planes per CTA (sized for the smaller resident set)

template <int K> using Int = std::integral_constant<int, K>;


// GEN = true : the scenario (SPEC.md §3) is generated in-kernel from the hash RNG straight into the bit
//              mask (no occupancy plane round trip); used by the batched env (reset and regeneration).
// GEN = false: the occupancy plane is an input (stateless operator), staged with one TMA bulk copy.
template <int WPR, bool GEN>
__global__ void __launch_bounds__(32, 16) flow_field_warp_kernel(FlowArgs a) {
    constexpr int NPS = nps_of(GEN);
    constexpr int NPLX = NPS + 2;               // + visited / free planes for the post-BFS phases
    constexpr int PVIS = NPS, PFREE = NPS + 1;
    constexpr int RPL = WPR;                    // rows per lane
    static_assert(WPR <= 3, "96 < G <= 128 is served by flow_field_il.cu");
    constexpr int P = 32 * WPR;                 // padded grid side
    constexpr int PLANE_WORDS = 32 * RPL * WPR;  // one bit-plane of the padded grid
    constexpr int CH = 2 * WPR;                 // 16-byte chunks per staged byte row
    constexpr bool SWZ = (CH & (CH - 1)) == 0;
    // cost-byte staging of one pass (32 rows x P bytes)
    constexpr bool INPLACE = false;
    constexpr int STAGE_WORDS = 8 * P;
    __shared__ __align__(128) uint32_t pl[NPLX * PLANE_WORDS];
    __shared__ __align__(16) uint32_t stage_buf[STAGE_WORDS];
    __shared__ __align__(8) uint64_t mbar;

    const int lane = threadIdx.x;
    const int G = a.G;
    const int count = a.count_ptr ? static_cast<int>(*a.count_ptr) : a.count;
    const uint32_t bar = static_cast<uint32_t>(__cvta_generic_to_shared(&mbar));
    const uint32_t pl_s = static_cast<uint32_t>(__cvta_generic_to_shared(pl));
    uint32_t *hi = a.hi_scratch + static_cast<size_t>(blockIdx.x) * (NPG * PLANE_WORDS);   // planes NPS..15
    const uint32_t neg1 = a.neg1;
    const uint32_t upm = lane == 0 ? 0u : a.one, dnm = lane == 31 ? 0u : a.one;   // boundary-lane masks (IMAD operands)
    uint32_t parity = 0;

    if (!GEN) {
        if (lane == 0) {
            mbar_init(bar, 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncwarp();
    }

    auto pidx = [&](int k, int r, int ln) { return ((k * RPL + r) * 32 + ln) * WPR; };
    auto plane_ptr = [&](int k, int r, int ln) -> uint32_t * {   // resident planes in shared memory, the others in L2
        return k < NPS ? &pl[pidx(k, r, ln)] : &hi[pidx(k - NPS, r, ln)];
    };

    // grids are handed out through a device counter (a.work), so warps that drew shallow grids take more of them
    for (int item = blockIdx.x;; item += gridDim.x) {
        if (a.work) {
            if (lane == 0) item = static_cast<int>(atomicAdd(a.work, 1u));
            item = __shfl_sync(FULL, item, 0);
        }
        if (item >= count) break;
        const uint32_t env = a.env_idx ? a.env_idx[item] : static_cast<uint32_t>(item);
        const size_t cells = static_cast<size_t>(G) * G;
        size_t plane;
        int gi, gj;
        uint32_t A[RPL][WPR], F[RPL][WPR], G0[RPL][WPR], G1[RPL][WPR];

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
            // rolled loop (code size): the free words go to their plane and are read back as rows below
#pragma unroll 1
            for (int rw = 0; rw < RPL * WPR; ++rw) {
                const int r = rw / WPR, w = rw - r * WPR;
                pl[pidx(PFREE, r, lane) + w] = scenario_free_word(key, lane * RPL + r, 32 * w, G, a.block_shift, a.p_thresh, sp);
            }
#pragma unroll
            for (int r = 0; r < RPL; ++r) Row<WPR>::ld(&pl[pidx(PFREE, r, lane)], A[r]);
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
            // ---- 1. TMA bulk copy of the occupancy plane into the plane storage (free before the BFS) -----------
            fence_proxy_async();  // earlier generic-proxy accesses to `pl` are ordered before the async write
            __syncwarp();
            if (lane == 0) {
                mbar_expect_tx(bar, static_cast<uint32_t>(cells));
                tma_bulk_g2s(pl_s, a.occ + plane * cells, static_cast<uint32_t>(cells), bar);
            }
            mbar_wait(bar, parity);
            parity ^= 1;

            // ---- 2. bytes -> free-cell bit mask: one byte per lane and one warp vote per 32 cells; the vote lands in
            //      the register of the lane that owns the row (static register indices, rolled over the owner lane) ----
            {
                const uint8_t *stage = reinterpret_cast<const uint8_t *>(pl) + lane;
                if (G == 32 * WPR) {
                    // exact fit: word (R, w) is the 32 bytes at (R * WPR + w) * 32; per word LDS.U8, compare, vote, select
#pragma unroll 1
                    for (int o = 0; o < 32; ++o) {
                        const bool mine = lane == o;
                        const uint8_t *src = stage + o * (RPL * WPR * 32);
#pragma unroll
                        for (int r = 0; r < RPL; ++r)
#pragma unroll
                            for (int w = 0; w < WPR; ++w) {
                                const uint32_t bits = __ballot_sync(FULL, src[(r * WPR + w) * 32] == 0);
                                if (mine) A[r][w] = bits;
                            }
                    }
                } else {
#pragma unroll 1
                    for (int o = 0; o < 32; ++o) {
                        const bool mine = lane == o;
#pragma unroll
                        for (int r = 0; r < RPL; ++r) {
                            const int R = o * RPL + r;
#pragma unroll
                            for (int w = 0; w < WPR; ++w) {
                                const int col = 32 * w + lane;
                                const bool fr = (R < G && col < G) ? stage[R * G + 32 * w] == 0 : false;
                                const uint32_t bits = __ballot_sync(FULL, fr);
                                if (mine) A[r][w] = bits;
                            }
                        }
                    }
                }
                __syncwarp();
            }
        }

        // ---- 3. the free mask goes to its plane; zero the resident Gray planes 2.. (0 and 1 live in registers) ----
        {
            uint32_t z[WPR];
#pragma unroll
            for (int w = 0; w < WPR; ++w) z[w] = 0;
#pragma unroll
            for (int r = 0; r < RPL; ++r) {
                if (!GEN) Row<WPR>::st(&pl[pidx(PFREE, r, lane)], A[r]);
#pragma unroll
                for (int k = 2; k < NPS; ++k) Row<WPR>::st(&pl[pidx(k, r, lane)], z);
            }
        }

        // ---- 4. bit-parallel wavefront -----------------------------------------------------------
#pragma unroll
        for (int r = 0; r < RPL; ++r)
#pragma unroll
            for (int w = 0; w < WPR; ++w) { G0[r][w] = 0; G1[r][w] = 0; }
        {
            // goal seeding with static register indices only (a conditional on (r, w) would turn A / F into local arrays)
            const bool ok = gi >= 0 && gj >= 0 && gi < G && gj < G && lane == gi / RPL;
            const int gr = gi % RPL, gw = gj >> 5;
            const uint32_t bit = ok ? (1u << (gj & 31)) : 0u;
#pragma unroll
            for (int r = 0; r < RPL; ++r)
#pragma unroll
                for (int w = 0; w < WPR; ++w) {
                    const uint32_t m = (r == gr && w == gw) ? (bit & A[r][w]) : 0u;
                    F[r][w] = m;
                    A[r][w] ^= m;
                }
        }
        // one BFS level; KSEL = 0 / 1: the Gray plane of this level is G0 / G1 (registers), 2: plane ctz(L) >= 2
        auto level = [&](auto ksel, uint32_t L) {
            constexpr int KSEL = decltype(ksel)::value;
            if constexpr (KSEL == 0) {
#pragma unroll
                for (int r = 0; r < RPL; ++r)
#pragma unroll
                    for (int w = 0; w < WPR; ++w) G0[r][w] ^= A[r][w];
            } else if constexpr (KSEL == 1) {
#pragma unroll
                for (int r = 0; r < RPL; ++r)
#pragma unroll
                    for (int w = 0; w < WPR; ++w) G1[r][w] ^= A[r][w];
            } else {
                const int k = __ffs(L) - 1;
                if (KSEL == 3 && k == 1) {
#pragma unroll
                    for (int r = 0; r < RPL; ++r)
#pragma unroll
                        for (int w = 0; w < WPR; ++w) G1[r][w] ^= A[r][w];
                } else if (k < NPS) {
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
                            uint32_t *p = &hi[pidx(k - NPS, r, lane) + w];
                            *p = first ? A[r][w] : (*p ^ A[r][w]);
                        }
                }
            }
            uint32_t upF[WPR], dnF[WPR];
#pragma unroll
            for (int w = 0; w < WPR; ++w) {
                if constexpr (GEN) {
                    // generated scenarios have occupied border rows (SPEC.md §3) and empty padding rows: what lane 0 / lane 31
                    // receive from themselves can only reach rows whose avail words are zero, so no boundary mask is needed
                    upF[w] = __shfl_up_sync(FULL, F[RPL - 1][w], 1);
                    dnF[w] = __shfl_down_sync(FULL, F[0][w], 1);
                } else {
                    upF[w] = mask_on_fma(__shfl_up_sync(FULL, F[RPL - 1][w], 1), upm);
                    dnF[w] = mask_on_fma(__shfl_down_sync(FULL, F[0][w], 1), dnm);
                }
            }
            uint32_t Nw[RPL][WPR];
#pragma unroll
            for (int r = 0; r < RPL; ++r)
#pragma unroll
                for (int w = 0; w < WPR; ++w) {
                    const uint32_t up = r == 0 ? upF[w] : F[r - 1][w];
                    const uint32_t dn = r == RPL - 1 ? dnF[w] : F[r + 1][w];
                    Nw[r][w] = (from_lo<WPR>(F[r], w) | from_hi<WPR>(F[r], w) | up | dn) & A[r][w];
                }
#pragma unroll
            for (int r = 0; r < RPL; ++r)
#pragma unroll
                for (int w = 0; w < WPR; ++w) {
                    A[r][w] = sub_on_fma(A[r][w], Nw[r][w], neg1);
                    F[r][w] = Nw[r][w];
                }
        };
        // Two levels of code (odd level: Gray plane 0, static; even level: plane ctz(L) >= 1, dynamic) executed twice per
        // convergence vote: half the instruction footprint of a four-level body (the loop showed 13 % no-instruction stalls).
        uint32_t L = 1;
        for (;; L += 4) {
#pragma unroll 1
            for (uint32_t h = 0; h < 4; h += 2) {
                level(Int<0>{}, L + h);
                level(Int<3>{}, L + h + 1);
            }
            uint32_t any = 0;                           // the frontier of the fourth level, OR-ed once per vote
#pragma unroll
            for (int r = 0; r < RPL; ++r)
#pragma unroll
                for (int w = 0; w < WPR; ++w) any |= F[r][w];
            if (!__any_sync(FULL, any != 0)) break;     // an empty frontier stays empty: test every fourth level
        }
#pragma unroll
        for (int r = 0; r < RPL; ++r) {
            Row<WPR>::st(&pl[pidx(0, r, lane)], G0[r]);
            Row<WPR>::st(&pl[pidx(1, r, lane)], G1[r]);
        }
        // level L+3 reached nothing, so the deepest level that reached a cell is <= L+2 (an over-estimate by at most
        // three levels only makes a few zero planes take part below)
        const uint32_t Lmax = L + 2;
        const int kmax = 32 - __clz(Lmax);                  // number of significant cost bits
#pragma unroll
        for (int r = 0; r < RPL; ++r) {
            uint32_t v[WPR], f[WPR];
            Row<WPR>::ld(&pl[pidx(PFREE, r, lane)], f);
#pragma unroll
            for (int w = 0; w < WPR; ++w) v[w] = f[w] & ~A[r][w];
            Row<WPR>::st(&pl[pidx(PVIS, r, lane)], v);
        }
        __syncwarp();

        // ---- 5. Gray -> binary, in place.  The (rarely more than two) planes that live in the global scratch are
        //      loaded up front as independent requests, so their L2 latency is paid once per row, not once per plane ----
#pragma unroll 1
        for (int r = 0; r < RPL; ++r) {
            uint32_t acc[WPR];
#pragma unroll
            for (int w = 0; w < WPR; ++w) acc[w] = 0;
            int k = kmax - 1;
#pragma unroll 1
            for (; k >= NPS + 4; --k) {          // deep maps only (depth >= 1024)
                uint32_t v[WPR];
                uint32_t *p = plane_ptr(k, r, lane);
                Row<WPR>::ld(p, v);
#pragma unroll
                for (int w = 0; w < WPR; ++w) { acc[w] ^= v[w]; v[w] = acc[w]; }
                Row<WPR>::st(p, v);
            }
            if (k >= NPS) {
                uint32_t g[4][WPR];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    if (NPS + j <= k) {
                        Row<WPR>::ld(plane_ptr(NPS + j, r, lane), g[j]);
                    } else {
#pragma unroll
                        for (int w = 0; w < WPR; ++w) g[j][w] = 0;
                    }
                }
#pragma unroll
                for (int j = 3; j >= 0; --j) {
                    if (NPS + j <= k) {
#pragma unroll
                        for (int w = 0; w < WPR; ++w) { acc[w] ^= g[j][w]; g[j][w] = acc[w]; }
                        Row<WPR>::st(plane_ptr(NPS + j, r, lane), g[j]);
                    }
                }
                k = NPS - 1;
            }
#pragma unroll 1
            for (; k >= 0; --k) {
                uint32_t v[WPR];
                uint32_t *p = &pl[pidx(k, r, lane)];
                Row<WPR>::ld(p, v);
#pragma unroll
                for (int w = 0; w < WPR; ++w) { acc[w] ^= v[w]; v[w] = acc[w]; }
                Row<WPR>::st(p, v);
            }
        }
        __syncwarp();

        // ---- 6. flow direction, bit-parallel, and the flow image ---------------------------------
        uint8_t *flow = a.flow + plane * cells;
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
            Row<WPR>::ld(&pl[pidx(PVIS, r, lane)], Vc);
            Row<WPR>::ld(&pl[pidx(PFREE, r, lane)], Fc);
            if (lu >= 0) {
                if (kmax > 1) Row<WPR>::ld(&pl[pidx(1, ru, lu)], b1u);
                if (kmax > 2) Row<WPR>::ld(&pl[pidx(2, ru, lu)], b2u);
                Row<WPR>::ld(&pl[pidx(PVIS, ru, lu)], Vu);
                Row<WPR>::ld(&pl[pidx(PFREE, ru, lu)], Fu);
            }
            if (ld < 32) {
                if (kmax > 1) Row<WPR>::ld(&pl[pidx(1, rd, ld)], b1d);
                if (kmax > 2) Row<WPR>::ld(&pl[pidx(2, rd, ld)], b2d);
                Row<WPR>::ld(&pl[pidx(PVIS, rd, ld)], Vd);
                Row<WPR>::ld(&pl[pidx(PFREE, rd, ld)], Fd);
            }
            const int R = lane * RPL + r;
#pragma unroll
            for (int w = 0; w < WPR; ++w) {
                const uint32_t own = Vc[w];
                const uint32_t t = b1c[w] ^ ~b0[w];   // bit 1 of (cost-1)
                const uint32_t u = b2c[w] ^ ~b1c[w];  // bit 2 of (cost-2)
                // orthogonal neighbours one level lower (codes 0 E, 2 N, 4 W, 6 S)
                const uint32_t lE = own & Vd[w] & ~(b1d[w] ^ t);
                const uint32_t lW = own & Vu[w] & ~(b1u[w] ^ t);
                const uint32_t lN = own & from_hi<WPR>(Vc, w) & ~(from_hi<WPR>(b1c, w) ^ t);
                const uint32_t lS = own & from_lo<WPR>(Vc, w) & ~(from_lo<WPR>(b1c, w) ^ t);
                // admissible diagonals two levels lower (codes 1 NE, 3 NW, 5 SW, 7 SE)
                const uint32_t fE = Fd[w], fW = Fu[w], fN = from_hi<WPR>(Fc, w), fS = from_lo<WPR>(Fc, w);
                const uint32_t lNE = own & from_hi<WPR>(Fd, w) & fE & fN & (from_hi<WPR>(b1d, w) ^ b1c[w]) & ~(from_hi<WPR>(b2d, w) ^ u);
                const uint32_t lNW = own & from_hi<WPR>(Fu, w) & fW & fN & (from_hi<WPR>(b1u, w) ^ b1c[w]) & ~(from_hi<WPR>(b2u, w) ^ u);
                const uint32_t lSW = own & from_lo<WPR>(Fu, w) & fW & fS & (from_lo<WPR>(b1u, w) ^ b1c[w]) & ~(from_lo<WPR>(b2u, w) ^ u);
                const uint32_t lSE = own & from_lo<WPR>(Fd, w) & fE & fS & (from_lo<WPR>(b1d, w) ^ b1c[w]) & ~(from_lo<WPR>(b2d, w) ^ u);
                const uint32_t anyD = lNE | lNW | lSW | lSE;
                const uint32_t m0 = (anyD & lNE) | (~anyD & lE);
                const uint32_t m1 = (anyD & lNW) | (~anyD & lN);
                const uint32_t m2 = (anyD & lSW) | (~anyD & lW);
                const uint32_t m3 = (anyD & lSE) | (~anyD & lS);
                // direction code bit-planes: d0 = diagonal, (d2 d1) = quadrant, d3 = none
                const uint32_t d1 = ~m0 & (m1 | (~m2 & m3));
                const uint32_t d2 = ~m0 & ~m1 & (m2 | m3);
                const uint32_t d3 = ~(m0 | m1 | m2 | m3);
                if (R < G && 32 * w < G) {
                    uint32_t out[8];
                    flow_bytes32(anyD, d1, d2, d3, ~Fc[w], out);
                    uint8_t *dst = flow + static_cast<size_t>(R) * G + 32 * w;
#pragma unroll
                    for (int c = 0; c < 2; ++c) {
                        const int col0 = 32 * w + 16 * c;
                        if (col0 + 16 <= G && (G & 15) == 0) {   // rows are 16-byte aligned only if G % 16 == 0
                            *reinterpret_cast<uint4 *>(dst + 16 * c) = make_uint4(out[4 * c], out[4 * c + 1], out[4 * c + 2], out[4 * c + 3]);
                        } else {
#pragma unroll
                            for (int n = 0; n < 4; ++n)
                                if (col0 + 4 * n < G) *reinterpret_cast<uint32_t *>(dst + 16 * c + 4 * n) = out[4 * c + n];
                        }
                    }
                }
            }
        }
        __syncwarp();

        // ---- 7. expand the integration field to int32 and store ---------------------------------
        if (a.cost) {
            int32_t *cost = a.cost + plane * cells;
            if (kmax <= NPL) {
                // fast path (depth < 256), one pass per row-of-the-lane r: the 8 plane words of the lane's row become 32
                // cost bytes per word (PRMT byte transposes + 8x8 bit transposes); the bytes of the 32 rows of this pass
                // are staged, XOR-swizzled, in the (k, r) slices they were read from, and every row is then widened by
                // the whole warp with one coalesced 16-byte store per lane
                uint8_t *pl8 = reinterpret_cast<uint8_t *>(pl);
#pragma unroll 1
                for (int r = 0; r < RPL; ++r) {
                    uint32_t Bk[NPL][WPR];
#pragma unroll
                    for (int k = 0; k < NPL; ++k) {
                        if (k < kmax) {
                            Row<WPR>::ld(plane_ptr(k, r, lane), Bk[k]);
                        } else {
#pragma unroll
                            for (int w = 0; w < WPR; ++w) Bk[k][w] = 0;
                        }
                    }
                    __syncwarp();
                    // logical staging buffer of this pass: 32 rows x P bytes
                    auto phys = [&](int o) -> uint8_t * {   // address of logical byte o (o multiple of 4)
                        if constexpr (INPLACE) {
                            constexpr int PB = PLANE_WORDS * 4, SL = 128 * WPR;    // bytes per plane / per (k, r) slice
                            if (o < PB) return pl8 + PFREE * PB + o;               // the free plane is dead by now
                            const int q = o - PB;
                            return pl8 + (((q / SL) * RPL + r) * 32 * WPR) * 4 + (q % SL);   // slice (k = q / SL, r) of planes 0..3
                        } else {
                            return reinterpret_cast<uint8_t *>(stage_buf) + o;
                        }
                    };
#pragma unroll
                    for (int w = 0; w < WPR; ++w) {
                        uint32_t tl[4], th[4];
                        bytes4x4(Bk[0][w], Bk[1][w], Bk[2][w], Bk[3][w], tl);
                        bytes4x4(Bk[4][w], Bk[5][w], Bk[6][w], Bk[7][w], th);
                        uint32_t out[8];
#pragma unroll
                        for (int b = 0; b < 4; ++b) {
                            uint32_t lo = tl[b], hb = th[b];
                            transpose8(lo, hb);          // byte j of hb:lo = cost of cell 8b+j
                            out[2 * b] = lo; out[2 * b + 1] = hb;
                        }
#pragma unroll
                        for (int c = 0; c < 2; ++c) {
                            const int chunk = 2 * w + c;
                            const int sw = SWZ ? (chunk ^ ((lane / (8 / CH)) % CH)) : chunk;
                            *reinterpret_cast<uint4 *>(phys(lane * P + 16 * sw)) =
                                make_uint4(out[4 * c], out[4 * c + 1], out[4 * c + 2], out[4 * c + 3]);
                        }
                    }
                    __syncwarp();
                    const int col = 4 * lane;
                    if (col < G) {
#pragma unroll 4
                        for (int i = 0; i < 32; ++i) {
                            const int R = i * RPL + r;
                            if (R >= G) break;
                            const int chunk = col >> 4;
                            const int sw = SWZ ? (chunk ^ ((i / (8 / CH)) % CH)) : chunk;
                            const uint32_t b4 = *reinterpret_cast<const uint32_t *>(phys(i * P + 16 * sw + (col & 15)));
                            uint32_t vb;
                            vb = pl[pidx(PVIS, r, i) + (col >> 5)] >> (col & 31);
                            int4 c;
                            c.x = (vb & 1u) ? static_cast<int>(b4 & 0xFFu) : COST_INF;
                            c.y = (vb & 2u) ? static_cast<int>((b4 >> 8) & 0xFFu) : COST_INF;
                            c.z = (vb & 4u) ? static_cast<int>((b4 >> 16) & 0xFFu) : COST_INF;
                            c.w = (vb & 8u) ? static_cast<int>(b4 >> 24) : COST_INF;
                            *reinterpret_cast<int4 *>(cost + static_cast<size_t>(R) * G + col) = c;
                        }
                    }
                    __syncwarp();
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
                        const uint32_t word = plane_ptr(k, r, lane)[w];
#pragma unroll
                        for (int n = 0; n < 8; ++n) lo[n] += spread4(word, n) << k;
                    }
#pragma unroll 1
                    for (int k = NPL; k < kmax; ++k) {
                        const uint32_t word = plane_ptr(k, r, lane)[w];
#pragma unroll
                        for (int n = 0; n < 8; ++n) hb[n] += spread4(word, n) << (k - NPL);
                    }
                    const uint32_t vis = pl[pidx(PVIS, r, lane) + w];
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
    }

    // the last CTA to finish re-arms the regeneration list for its next use
    if (a.ticket && lane == 0) flow_launch_epilogue(a);
}

}  // namespace

bool flow_field_large_supported(int G);
int flow_field_large_max_grid(int G);
size_t flow_field_large_scratch_words(int G);
bool flow_field_rows_usable(int G);

// FFMP_FLOW_ROWS=1: grids up to 128 with G % 32 == 0 also use the CTA-per-grid row kernel of flow_field_large.cu (lower
// latency per grid and a smaller footprint for the env's background regeneration, lower batch throughput; measured in
// profiles/r01d_rows_small.txt).  Scratch sizes cover both kernels, so the switch may change between calls.
static bool rows_for_small() {
    const char *e = std::getenv("FFMP_FLOW_ROWS");
    return e && std::atoi(e) != 0;
}
cudaError_t launch_flow_field_large(const FlowArgs &a, int grid, cudaStream_t st);
cudaError_t launch_flow_field_il(const FlowArgs &a, int grid, cudaStream_t st);      // flow_field_il.cu: 96 < G <= 128
cudaError_t launch_flow_field_wide(const FlowArgs &a, int grid, cudaStream_t st);    // flow_field_wide.cu: 384 < G <= 512
bool flow_field_wide_supported(int G);
size_t flow_field_wide_scratch_words();
int flow_field_wide_max_grid();

bool flow_field_supported(int G) { return (G >= 16 && G <= 128 && (G % 4) == 0) || flow_field_large_supported(G); }

size_t flow_field_scratch_words(int G) {
    if (flow_field_wide_supported(G)) return flow_field_wide_scratch_words();
    if (G > 128) return flow_field_large_scratch_words(G);
    if (flow_field_rows_usable(G)) return flow_field_large_scratch_words(G);     // >= the warp kernel's need
    // planes NPS..15 of the padded grid ((G+31)/32*32)^2 live in the per-CTA scratch (L2 resident)
    const int wpr = (G + 31) / 32;
    return static_cast<size_t>(NPG) * 32 * wpr * wpr;
}

int flow_field_il_ctas_per_sm();
int flow_field_max_grid(int G) {
    if (flow_field_wide_supported(G)) return flow_field_wide_max_grid();
    if (G > 128) return flow_field_large_max_grid(G);
    const int wpr = (G + 31) / 32;
    if (wpr == 4) return 148 * flow_field_il_ctas_per_sm();
    const int smem = (nps_of(false) + 2) * 32 * wpr * wpr * 4 + (wpr == 4 ? 16 : 32 * 32 * wpr) + 1024 + 16;
    int per_sm = (227 * 1024) / smem;
    if (per_sm > 16) per_sm = 16;            // __launch_bounds__(32, 16): 128 registers per thread
    return 148 * per_sm;
}

// FlowArgs.all_slots is understood by the interleaved-layout kernels (flow_field_il.cu) only
bool flow_field_takes_all_slots(int G) {
    return G > 96 && G <= 128 && !flow_field_wide_supported(G) && !(flow_field_rows_usable(G) && rows_for_small());
}

cudaError_t launch_flow_field(const FlowArgs &a_in, int grid, cudaStream_t st) {
    if (a_in.all_slots && !flow_field_takes_all_slots(a_in.G)) return cudaErrorInvalidValue;
    if (grid <= 0) return cudaSuccess;
    if (flow_field_wide_supported(a_in.G)) return launch_flow_field_wide(a_in, grid, st);
    if (a_in.G > 128 || (flow_field_rows_usable(a_in.G) && rows_for_small())) return launch_flow_field_large(a_in, grid, st);
    FlowArgs a = a_in;
    a.neg1 = 0xFFFFFFFFu;
    a.one = 1u;
    const int wpr = (a.G + 31) / 32;
    if (a.work && !a.ticket) return cudaErrorInvalidValue;   // the work counter is re-armed by the ticket holder
    if (wpr == 4) return launch_flow_field_il(a, grid, st);
    if (a.generate) {
        switch (wpr) {
        case 1: flow_field_warp_kernel<1, true><<<grid, 32, 0, st>>>(a); break;
        case 2: flow_field_warp_kernel<2, true><<<grid, 32, 0, st>>>(a); break;
        case 3: flow_field_warp_kernel<3, true><<<grid, 32, 0, st>>>(a); break;
        default: return cudaErrorInvalidValue;
        }
    } else {
        switch (wpr) {
        case 1: flow_field_warp_kernel<1, false><<<grid, 32, 0, st>>>(a); break;
        case 2: flow_field_warp_kernel<2, false><<<grid, 32, 0, st>>>(a); break;
        case 3: flow_field_warp_kernel<3, false><<<grid, 32, 0, st>>>(a); break;
        default: return cudaErrorInvalidValue;
        }
    }
    return cudaGetLastError();
}

}  // namespace ffmp
