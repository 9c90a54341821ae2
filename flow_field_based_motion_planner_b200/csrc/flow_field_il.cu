// flow_field_il.cu — SPEC.md §4/§5 for 96 < G <= 128: integration field (F1), flow direction (F2), flow image.
// Replaces the external /bev/* flow-image ROS node (/root/reference/src/train.py:84,116-121).
//
// One WARP per grid, no block barrier.  Lane l owns rows 4l .. 4l+3 of the free / frontier masks as 16 + 16 registers in the
// column-interleaved layout of flow_rowops.cuh (word w of a row = the columns c % 4 == w), so the horizontal neighbours of the
// wavefront step are the neighbouring words themselves (2 shifts per row per level) and the vertical ones the neighbouring row
// registers (8 shuffles per level for the lane-boundary rows).
//
// What is new against the round-1 kernel (flow_field.cu keeps serving G <= 96):
//   * cost bit 0 is never recorded: on a 4-connected unit-cost field it is the checkerboard colour (i + j + gi + gj) & 1.  The
//     Gray-code bit-planes record M = level >> 1, so a plane is touched on every other level only and plane ctz(M) at that;
//   * a plane update is an ADD, not an XOR: the sets of cells that are still unreached at successive toggles of one plane are
//     nested, so the XOR of them equals their alternating sum, which is an IMAD with a +-1 multiplier on the FMA pipe.  With
//     avail -= new, the lane-boundary masks and the wrap-around left shifts there as well, the level loop issues 36 ALU-pipe
//     and 36 FMA-pipe instructions per level instead of 56 + 20 (both pipes issue every other cycle per scheduler);
//   * the occupancy plane is packed by the lane that owns the row: 16-byte shared-memory reads (rotated by the lane so that the
//     512-byte lane stride stays conflict-free), a SWAR non-zero test, and IMADs to deposit the nibbles — no warp votes;
//   * cost bytes come from one 32 x 32 bit-matrix transpose per row (PRMT byte stages + three bit-select stages) and are widened
//     with ONE PRMT per cell (the "not reached" flag travels through the transpose as bit 0 of the byte, which makes INF a byte
//     pattern instead of a per-cell select); flow bytes come from a 16 x 16 x 2 transpose of 4-bit selector planes and ONE
//     PRMT table look-up per four cells.
// Algorithmic HBM bytes: 6 B/cell (1 occupancy read + 4 cost write + 1 flow write).
#include <cstdlib>
#include <type_traits>

#include "flow_bits.cuh"
#include "flow_rowops.cuh"

namespace ffmp {

namespace {

constexpr int IL_NPS = 6;                  // resident Gray planes: bits 0..5 of M = level >> 1 (levels < 128)
constexpr int IL_NPG = 12;                 // planes of the per-CTA global scratch: bits 6..17 of M
constexpr int IL_PVIS = IL_NPS, IL_PFREE = IL_NPS + 1, IL_NPLX = IL_NPS + 2;
constexpr int IL_PW = 32 * 4 * 4;          // words of one bit-plane of the padded 128 x 128 grid
constexpr int IL_STAGE_WORDS = 32 * 32;    // output staging of one pass: 32 rows x 128 bytes
constexpr int IL_WARPS_PER_SM = 10;        // 20.5 KB of shared memory per warp (+ 1 KB per CTA reserved by the driver)

__device__ __forceinline__ uint32_t mad_u32(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t r;
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
    return r;
}

__device__ __forceinline__ int pidx(int k, int r, int ln) { return ((k * 4 + r) * 32 + ln) * 4; }
__device__ __forceinline__ void ld4(const uint32_t *p, uint32_t (&v)[4]) {
    const uint4 t = *reinterpret_cast<const uint4 *>(p);
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
}
__device__ __forceinline__ void st4(uint32_t *p, const uint32_t (&v)[4]) { *reinterpret_cast<uint4 *>(p) = make_uint4(v[0], v[1], v[2], v[3]); }

// One pass of the output phase: row R = 4 * lane + r of the grid for every lane of the calling warp (32 rows at a time).  Flow
// direction -> flow bytes; cost planes -> int32.  Both outputs go through the warp's swizzled staging buffer so that a warp
// store instruction writes whole 128-byte lines (a lane that stores its own row 16 bytes at a time costs the LSU 32 wavefronts
// per instruction).  `pl` = the binary cost planes / reached / free planes in shared memory, `hi` = the L2 scratch planes.
__device__ __forceinline__ void il_emit_rows(uint32_t *pl, const uint32_t *hi, uint32_t *stage, const int r, const int lane,
                                             const int G, const int gi, const int gj, const int kmax, uint8_t *flow,
                                             int32_t *cost) {
    // staging: row i (one per lane) is 8 chunks of 16 bytes; chunk c sits at position c ^ (i & 7), so that both the row-wise
    // 16-byte writes of 8 consecutive lanes and the reads along a row are free of bank conflicts
    auto stage_chunk = [&](int i, int c) -> uint32_t * { return &stage[i * 32 + ((c ^ (i & 7)) << 2)]; };
    const int R = lane * 4 + r;
    const int lu = r == 0 ? lane - 1 : lane, ru = r == 0 ? 3 : r - 1;          // row R-1
    const int ldn = r == 3 ? lane + 1 : lane, rd = r == 3 ? 0 : r + 1;         // row R+1
    const uint32_t par0 = static_cast<uint32_t>(R + gi + gj) & 1u;
    uint32_t bk6[4] = {0u, 0u, 0u, 0u};
    if (cost && kmax == 7) ld4(&hi[pidx(0, r, lane)], bk6);                    // cost bit 7 (L2): requested early
    rowops::RowIn in;
    ld4(&pl[pidx(0, r, lane)], in.b1c); ld4(&pl[pidx(1, r, lane)], in.b2c);
    ld4(&pl[pidx(IL_PVIS, r, lane)], in.Vc); ld4(&pl[pidx(IL_PFREE, r, lane)], in.Fc);
#pragma unroll
    for (int w = 0; w < 4; ++w) {
        in.b1u[w] = in.b2u[w] = in.Vu[w] = in.Fu[w] = 0u;
        in.b1d[w] = in.b2d[w] = in.Vd[w] = in.Fd[w] = 0u;
    }
    // a free neighbour of a reached cell is reached: the neighbour rows' reached masks stand in for their free masks
    if (lu >= 0) {
        ld4(&pl[pidx(0, ru, lu)], in.b1u); ld4(&pl[pidx(1, ru, lu)], in.b2u);
        ld4(&pl[pidx(IL_PVIS, ru, lu)], in.Vu);
#pragma unroll
        for (int w = 0; w < 4; ++w) in.Fu[w] = in.Vu[w];
    }
    if (ldn < 32) {
        ld4(&pl[pidx(0, rd, ldn)], in.b1d); ld4(&pl[pidx(1, rd, ldn)], in.b2d);
        ld4(&pl[pidx(IL_PVIS, rd, ldn)], in.Vd);
#pragma unroll
        for (int w = 0; w < 4; ++w) in.Fd[w] = in.Vd[w];
    }
    {
        uint32_t n[4][4], fw[32];
        rowops::direction_nibbles(in, par0, n);
        rowops::flow_row_words(n, fw);
#pragma unroll
        for (int c = 0; c < 8; ++c)
            *reinterpret_cast<uint4 *>(stage_chunk(lane, c)) = make_uint4(fw[4 * c], fw[4 * c + 1], fw[4 * c + 2], fw[4 * c + 3]);
    }
    __syncwarp();
    if ((G & 15) == 0) {
        // 8 lanes per row, 4 rows per instruction: four whole lines
        const int c = lane & 7;
#pragma unroll
        for (int i0 = 0; i0 < 32; i0 += 4) {
            const int i = i0 + (lane >> 3), Ri = 4 * i + r;
            const uint4 t = *reinterpret_cast<const uint4 *>(stage_chunk(i, c));
            if (Ri < G && 16 * c < G) *reinterpret_cast<uint4 *>(flow + static_cast<size_t>(Ri) * G + 16 * c) = t;
        }
    } else {
#pragma unroll 4
        for (int i = 0; i < 32; ++i) {
            const int Ri = 4 * i + r;
            const uint32_t t = stage_chunk(i, lane >> 2)[lane & 3];
            if (Ri < G && 4 * lane < G) *reinterpret_cast<uint32_t *>(flow + static_cast<size_t>(Ri) * G + 4 * lane) = t;
        }
    }
    __syncwarp();
    if (cost) {
        if (kmax <= 7) {
            // depth < 256: one 32 x 32 bit transpose per row; the row's 32 words (4 cells each) are staged, and lane b
            // widens word b of every row with one PRMT per cell: a store instruction writes 512 contiguous bytes
            uint32_t x[32];
#pragma unroll
            for (int w = 0; w < 4; ++w) x[8 * w] = ~in.Vc[w];
#pragma unroll
            for (int k = 0; k < 7; ++k) {
                uint32_t bk[4];
                if (k < IL_NPS) ld4(&pl[pidx(k, r, lane)], bk);
#pragma unroll
                for (int w = 0; w < 4; ++w) x[8 * w + 1 + k] = (k < IL_NPS ? bk[w] : bk6[w]) | ~in.Vc[w];
            }
            rowops::transpose32(x);
#pragma unroll
            for (int c = 0; c < 8; ++c)
                *reinterpret_cast<uint4 *>(stage_chunk(lane, c)) = make_uint4(x[4 * c], x[4 * c + 1], x[4 * c + 2], x[4 * c + 3]);
            __syncwarp();
            const uint32_t parw = ((static_cast<uint32_t>(r + gi + gj)) & 1u) ? 0x00010001u : 0x01000100u;
#pragma unroll 8
            for (int i = 0; i < 32; ++i) {
                const int Ri = 4 * i + r;
                const uint32_t T = stage_chunk(i, lane >> 2)[lane & 3];
                const rowops::Int4 c4 = rowops::widen_cost4(T, parw);
                if (Ri < G && 4 * lane < G)
                    *reinterpret_cast<int4 *>(cost + static_cast<size_t>(Ri) * G + 4 * lane) = make_int4(c4.x, c4.y, c4.z, c4.w);
            }
            __syncwarp();
        } else if (R < G) {
            // deep maps (>= 256 levels: mazes): cell by cell from the planes
            int32_t *dst = cost + static_cast<size_t>(R) * G;
#pragma unroll 1
            for (int b = 0; 4 * b < G; ++b) {
                uint32_t v[4];
#pragma unroll
                for (int w = 0; w < 4; ++w) v[w] = (par0 ^ static_cast<uint32_t>(w)) & 1u;
#pragma unroll 1
                for (int k = 0; k < kmax; ++k) {
                    uint32_t t[4];
                    ld4(k < IL_NPS ? &pl[pidx(k, r, lane)] : &hi[pidx(k - IL_NPS, r, lane)], t);
#pragma unroll
                    for (int w = 0; w < 4; ++w) v[w] |= ((t[w] >> b) & 1u) << (k + 1);
                }
                int4 c;
                c.x = (in.Vc[0] >> b) & 1u ? static_cast<int>(v[0]) : COST_INF;
                c.y = (in.Vc[1] >> b) & 1u ? static_cast<int>(v[1]) : COST_INF;
                c.z = (in.Vc[2] >> b) & 1u ? static_cast<int>(v[2]) : COST_INF;
                c.w = (in.Vc[3] >> b) & 1u ? static_cast<int>(v[3]) : COST_INF;
                *reinterpret_cast<int4 *>(dst + 4 * b) = c;
            }
        }
    }
}

// GEN = true : the scenario (SPEC.md §3) is generated in-kernel from the hash RNG straight into the bit mask (batched env:
//              reset and background regeneration).
// GEN = false: the occupancy plane is an input (stateless operator), staged with one TMA bulk copy.
template <bool GEN>
__global__ void __launch_bounds__(32, IL_WARPS_PER_SM) flow_field_il_kernel(FlowArgs a) {
    __shared__ __align__(128) uint32_t pl[IL_NPLX * IL_PW];      // 16 KB: Gray planes 0..5, reached, free (+ the TMA staging)
    __shared__ __align__(128) uint32_t stage[IL_STAGE_WORDS];    // 4 KB: a pass's output rows, chunk-swizzled
    __shared__ __align__(8) uint64_t mbar;

    const int lane = threadIdx.x;
    const int G = a.G;
    const int count = a.count_ptr ? static_cast<int>(*a.count_ptr) : a.count;
    const uint32_t bar = static_cast<uint32_t>(__cvta_generic_to_shared(&mbar));
    const uint32_t pl_s = static_cast<uint32_t>(__cvta_generic_to_shared(pl));
    uint32_t *hi = a.hi_scratch + static_cast<size_t>(blockIdx.x) * (IL_NPG * IL_PW);
    const uint32_t neg1 = a.neg1, one = a.one, two = a.one + a.one;      // opaque IMAD multipliers (set by the launcher)
    const uint32_t upm = lane == 0 ? 0u : one, dnm = lane == 31 ? 0u : one;
    uint32_t parity = 0;

    if (!GEN) {
        if (lane == 0) {
            mbar_init(bar, 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncwarp();
    }

    for (int item = blockIdx.x;; item += gridDim.x) {
        if (a.work) {
            if (lane == 0) item = static_cast<int>(atomicAdd(a.work, 1u));
            item = __shfl_sync(FULL, item, 0);
        }
        if (item >= count) break;
        const uint32_t env = a.all_slots ? static_cast<uint32_t>(item % a.N)
                                         : (a.env_idx ? a.env_idx[item] : static_cast<uint32_t>(item));
        const size_t cells = static_cast<size_t>(G) * G;
        size_t plane;
        int gi, gj;
        uint32_t A[4][4], F[4][4], P0[4][4];

        if (GEN) {
            // ---- 1g. scenario parameters (lane 0) and the free-cell mask straight from the hash ----
            const uint32_t episode = a.all_slots ? a.episode_const + static_cast<uint32_t>(item / a.N)
                                                 : (a.episode ? a.episode[item] : a.episode_const);
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
#pragma unroll 1
            for (int r = 0; r < 4; ++r) {      // rolled (code size): the rows go to the free plane and are read back below
                uint32_t fr[4];
                scenario_free_row_il(key, lane * 4 + r, G, a.block_shift, a.p_thresh, sp, fr);
                st4(&pl[pidx(IL_PFREE, r, lane)], fr);
            }
#pragma unroll
            for (int r = 0; r < 4; ++r) ld4(&pl[pidx(IL_PFREE, r, lane)], A[r]);
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
            // ---- 1. TMA bulk copy of the occupancy plane into the plane storage (free before the BFS) ----
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) {
                mbar_expect_tx(bar, static_cast<uint32_t>(cells));
                tma_bulk_g2s(pl_s, a.occ + plane * cells, static_cast<uint32_t>(cells), bar);
            }
            mbar_wait(bar, parity);
            parity ^= 1;
            // ---- 2. bytes -> the lane's 16 free-mask words ----
            const uint8_t *bytes = reinterpret_cast<const uint8_t *>(pl);
            if (G == 128) {
                // the lane packs its own rows: chunk (q + lane) % 8 of the row first, so that the eight lanes of a 16-byte
                // shared-memory phase hit eight different bank groups despite their 512-byte stride.  Rolled over the four
                // rows (code size); the packed rows are parked in the output staging buffer, which is idle here
#pragma unroll 1
                for (int r = 0; r < 4; ++r) {
                    uint32_t o[4] = {0u, 0u, 0u, 0u};
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        const int qq = (q + lane) & 7;
                        const uint4 x = *reinterpret_cast<const uint4 *>(bytes + (4 * lane + r) * 128 + 16 * qq);
                        const uint32_t z = rowops::occupied_nibbles(x.x, x.y, x.z, x.w);
                        const uint32_t sh = 1u << (4 * qq);
#pragma unroll
                        for (int w = 0; w < 4; ++w) o[w] = mad_u32(rowops::prmt(z, 0u, 0x4440u + w), sh, o[w]);
                    }
#pragma unroll
                    for (int w = 0; w < 4; ++w) o[w] = ~o[w];
                    st4(&stage[(r * 32 + lane) * 4], o);
                }
#pragma unroll
                for (int r = 0; r < 4; ++r) ld4(&stage[(r * 32 + lane) * 4], A[r]);
            } else {
                // 96 < G < 128 (rows are not 16-byte aligned): one byte per lane and one warp vote per word; lane l votes for
                // column 4l + w, so the vote IS interleaved word w of the row
#pragma unroll 1
                for (int o = 0; o < 32; ++o) {
                    const bool mine = lane == o;
#pragma unroll
                    for (int r = 0; r < 4; ++r) {
                        const int R = o * 4 + r;
#pragma unroll
                        for (int w = 0; w < 4; ++w) {
                            const int col = 4 * lane + w;
                            const bool fr = (R < G && col < G) ? bytes[R * G + col] == 0 : false;
                            const uint32_t bits = __ballot_sync(FULL, fr);
                            if (mine) A[r][w] = bits;
                        }
                    }
                }
            }
            __syncwarp();          // every lane is done with the staged bytes before the planes are written
        }

        // ---- 3. the free mask goes to its plane; zero the resident Gray planes 1.. (plane 0 lives in registers) ----
        {
            const uint32_t z[4] = {0u, 0u, 0u, 0u};
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                if (!GEN) st4(&pl[pidx(IL_PFREE, r, lane)], A[r]);
#pragma unroll
                for (int k = 1; k < IL_NPS; ++k) st4(&pl[pidx(k, r, lane)], z);
            }
        }

        // ---- 4. bit-parallel wavefront ----
        {
            // goal seeding with static register indices only
            const bool ok = gi >= 0 && gj >= 0 && gi < G && gj < G && lane == (gi >> 2);
            const int gr = gi & 3, gw = gj & 3;
            const uint32_t bit = ok ? (1u << (gj >> 2)) : 0u;
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int w = 0; w < 4; ++w) {
                    const uint32_t m = (r == gr && w == gw) ? (bit & A[r][w]) : 0u;
                    F[r][w] = m;
                    A[r][w] ^= m;
                    P0[r][w] = 0u;
                }
        }
        uint32_t L = 1;
        // MASKED = false when rows 0 and 127 have no free cell (every generated scenario, SPEC.md §3): what lanes 0 / 31 receive
        // from themselves in the shuffles can then only reach rows whose avail words are zero, and the eight boundary-lane
        // masks per level are not needed.
        auto wavefront = [&](auto masked) {
            constexpr bool MASKED = decltype(masked)::value;
            // upF / dnF: row 4l-1 / 4l+4 of the frontier, requested one level ahead (the shuffles of level n+1 are issued as
            // soon as the lane-boundary rows of level n are known and complete while the inner rows are computed)
            uint32_t upF[4], dnF[4];
            auto exchange = [&](const uint32_t (&top)[4], const uint32_t (&bottom)[4]) {
#pragma unroll
                for (int w = 0; w < 4; ++w) {
                    upF[w] = __shfl_up_sync(FULL, bottom[w], 1);
                    dnF[w] = __shfl_down_sync(FULL, top[w], 1);
                }
            };
            auto lo = [&](const uint32_t (&x)[4], int w) { return w > 0 ? x[w - 1] : mask_on_fma(x[3], two); };   // column - 1
            auto hh = [&](const uint32_t (&x)[4], int w) { return w < 3 ? x[w + 1] : x[0] >> 1; };                 // column + 1
            auto step = [&]() {       // one BFS level: new = (W | E | N | S of the frontier) & avail
                uint32_t N0[4], N1[4], N2[4], N3[4], u[4], d[4];
#pragma unroll
                for (int w = 0; w < 4; ++w) {
                    u[w] = MASKED ? mask_on_fma(upF[w], upm) : upF[w];
                    d[w] = MASKED ? mask_on_fma(dnF[w], dnm) : dnF[w];
                }
#pragma unroll
                for (int w = 0; w < 4; ++w) {
                    N0[w] = (lo(F[0], w) | hh(F[0], w) | u[w] | F[1][w]) & A[0][w];
                    N3[w] = (lo(F[3], w) | hh(F[3], w) | F[2][w] | d[w]) & A[3][w];
                }
                exchange(N0, N3);
#pragma unroll
                for (int w = 0; w < 4; ++w) {
                    N1[w] = (lo(F[1], w) | hh(F[1], w) | F[0][w] | F[2][w]) & A[1][w];
                    N2[w] = (lo(F[2], w) | hh(F[2], w) | F[1][w] | F[3][w]) & A[2][w];
                }
#pragma unroll
                for (int w = 0; w < 4; ++w) {
                    A[0][w] = sub_on_fma(A[0][w], N0[w], neg1); F[0][w] = N0[w];
                    A[1][w] = sub_on_fma(A[1][w], N1[w], neg1); F[1][w] = N1[w];
                    A[2][w] = sub_on_fma(A[2][w], N2[w], neg1); F[2][w] = N2[w];
                    A[3][w] = sub_on_fma(A[3][w], N3[w], neg1); F[3][w] = N3[w];
                }
            };
            exchange(F[0], F[3]);
            uint32_t s0 = one;                // sign of the next toggle of plane 0
            for (;; L += 4) {
                // levels L .. L+3 (L = 4t + 1).  Level L+1: M = 2t+1, plane 0 (registers); level L+3: M = 2t+2, plane ctz(M) >= 1
                const uint32_t M = (L + 3) >> 1;
                const int k = __ffs(M) - 1;
                const uint32_t sk = ((M >> (k + 1)) & 1u) ? neg1 : one;
                const bool resident = k < IL_NPS;
                uint32_t *ps = &pl[pidx(resident ? k : 0, 0, lane)];                    // shared-memory plane (levels < 128)
                uint32_t *pg = &hi[pidx(resident ? 0 : k - IL_NPS, 0, lane)];           // L2 scratch plane
                const bool first = M == (1u << k);                   // the scratch is not zeroed: its first toggle stores
                step();
#pragma unroll
                for (int r = 0; r < 4; ++r)
#pragma unroll
                    for (int w = 0; w < 4; ++w) P0[r][w] = mad_u32(A[r][w], s0, P0[r][w]);
                s0 = 0u - s0;
                step();
                // the plane of level L+3 is requested one level ahead of its update
                uint32_t v[4][4];
                if (resident) {
#pragma unroll
                    for (int r = 0; r < 4; ++r) ld4(ps + r * 128, v[r]);
                } else {
#pragma unroll
                    for (int r = 0; r < 4; ++r) {
                        v[r][0] = v[r][1] = v[r][2] = v[r][3] = 0u;
                        if (!first) ld4(pg + r * 128, v[r]);
                    }
                }
                step();
#pragma unroll
                for (int r = 0; r < 4; ++r)
#pragma unroll
                    for (int w = 0; w < 4; ++w) v[r][w] = mad_u32(A[r][w], sk, v[r][w]);
                if (resident) {
#pragma unroll
                    for (int r = 0; r < 4; ++r) st4(ps + r * 128, v[r]);
                } else {
#pragma unroll
                    for (int r = 0; r < 4; ++r) st4(pg + r * 128, v[r]);
                }
                step();
                uint32_t any = 0;
#pragma unroll
                for (int r = 0; r < 4; ++r)
#pragma unroll
                    for (int w = 0; w < 4; ++w) any |= F[r][w];
                if (!__any_sync(FULL, any != 0)) break;      // an empty frontier stays empty: test every fourth level
            }
        };
        if constexpr (GEN) {
            wavefront(std::false_type{});
        } else {
            uint32_t edge = 0;
#pragma unroll
            for (int w = 0; w < 4; ++w) edge |= (lane == 0 ? A[0][w] : 0u) | (lane == 31 ? A[3][w] : 0u);
            if (__any_sync(FULL, edge != 0)) wavefront(std::true_type{});
            else wavefront(std::false_type{});
        }
        // level L+3 reached nothing: the deepest level is <= L+2 (an over-estimate only makes zero planes take part)
        const uint32_t Mmax = (L + 2) >> 1;
        const int kmax = 32 - __clz(Mmax);               // significant bits of M; cost < 2^(kmax+1)
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            uint32_t v[4], f[4];
            ld4(&pl[pidx(IL_PFREE, r, lane)], f);
#pragma unroll
            for (int w = 0; w < 4; ++w) v[w] = f[w] & ~A[r][w];
            st4(&pl[pidx(IL_PVIS, r, lane)], v);
            st4(&pl[pidx(0, r, lane)], P0[r]);
        }

        // ---- 5. Gray -> binary, in place (plane k becomes cost bit k + 1) ----
        if (kmax <= 7) {
            // the common case: at most one plane lives in the L2 scratch, and as the top plane its binary form is itself;
            // the four rows' requests are issued together
            uint32_t top[4][4];
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                top[r][0] = top[r][1] = top[r][2] = top[r][3] = 0u;
                if (kmax == 7) ld4(&hi[pidx(0, r, lane)], top[r]);
            }
#pragma unroll
            for (int r = 0; r < 4; ++r) {
#pragma unroll
                for (int k = IL_NPS - 1; k >= 0; --k) {
                    uint32_t v[4];
                    uint32_t *p = &pl[pidx(k, r, lane)];
                    ld4(p, v);
#pragma unroll
                    for (int w = 0; w < 4; ++w) top[r][w] ^= v[w];
                    st4(p, top[r]);
                }
            }
        } else {
#pragma unroll 1
            for (int r = 0; r < 4; ++r) {
                uint32_t acc[4] = {0u, 0u, 0u, 0u};
#pragma unroll 1
                for (int k = kmax - 1; k >= 0; --k) {
                    uint32_t v[4];
                    uint32_t *p = k < IL_NPS ? &pl[pidx(k, r, lane)] : &hi[pidx(k - IL_NPS, r, lane)];
                    ld4(p, v);
#pragma unroll
                    for (int w = 0; w < 4; ++w) acc[w] ^= v[w];
                    st4(p, acc);
                }
            }
        }
        __syncwarp();

        // ---- 6. per row-of-the-lane r (32 rows of the grid at a time): flow direction -> flow bytes; cost planes -> int32.
        //      Both outputs go through the swizzled staging buffer so that a warp store instruction writes whole 128-byte
        //      lines (a lane that stores its own row 16 bytes at a time costs the LSU 32 wavefronts per instruction) ----
        uint8_t *flow = a.flow + plane * cells;
        int32_t *cost = a.cost ? a.cost + plane * cells : nullptr;
#pragma unroll 1
        for (int r = 0; r < 4; ++r) {
            il_emit_rows(pl, hi, stage, r, lane, G, gi, gj, kmax, flow, cost);
        }
        __syncwarp();
    }

    // the last CTA to finish re-arms the regeneration list for its next use
    if (a.ticket && lane == 0) flow_launch_epilogue(a);
}


// ---- latency-oriented variant: FOUR warps per grid (background regeneration, small resets) ---------------------------------
// The warp-per-grid kernel above is the throughput form: no barrier, 16 + 16 mask registers per lane, but one grid takes
// ~190 levels x 72 instructions on ONE warp — 35-70 us when the SM is shared with step kernels — and the last regeneration of
// a rollout (or the reset of a single env, the reference's own use) waits for exactly that.  Here thread (warp r, lane l) owns
// row 4l + r only (4 + 4 mask registers): a level is 8 LOP3 per thread; the rows above / below belong to other warps and are
// exchanged through a double-buffered 2 KB shared-memory array (slot 32r + l: conflict-free 16-byte accesses), ordered by one
// block barrier per level; convergence is voted every fourth level by the barrier itself.  The row <-> thread mapping is the
// one the output phase wants (pass r = rows 4l + r): each warp emits its own pass through its own staging buffer, so the
// scenario rows, the Gray -> binary conversion and the output run four-wide as well.  Scenario generation only (GEN): rows 0
// and G-1.. of a generated map are walls, so what the first / last row reads past the grid never matters.
// Same Gray planes, same row arithmetic and same bytes out as the warp kernel (tests compare both against the oracle).
constexpr int QUAD_CTAS_PER_SM = 3;

__global__ void __launch_bounds__(128, 4) flow_field_quad_kernel(FlowArgs a) {
    __shared__ __align__(128) uint32_t pl[IL_NPLX * IL_PW];          // 16 KB: Gray planes 0..5, reached, free
    __shared__ __align__(128) uint32_t stage[4][IL_STAGE_WORDS];     // 16 KB: one output staging buffer per warp
    __shared__ __align__(16) uint32_t xch[2][128 * 4];               // 4 KB: the frontier rows of the current / next level
    __shared__ int s_item;

    const int lane = threadIdx.x & 31, r = threadIdx.x >> 5;
    const int G = a.G;
    const int count = a.count_ptr ? static_cast<int>(*a.count_ptr) : a.count;
    uint32_t *hi = a.hi_scratch + static_cast<size_t>(blockIdx.x) * (IL_NPG * IL_PW);
    const uint32_t neg1 = a.neg1, one = a.one, two = a.one + a.one;
    const int slot = (r * 32 + lane) * 4;
    const int up_slot = (r > 0 ? (r - 1) * 32 + lane : 96 + lane - 1) * 4;      // row 4l + r - 1 (lane 0 of warp 0: any row)
    const int dn_slot = (r < 3 ? (r + 1) * 32 + lane : lane + 1) * 4;           // row 4l + r + 1 (lane 31 of warp 3: any row)

    for (int item = blockIdx.x;; item += gridDim.x) {
        __syncthreads();            // the previous grid's planes have been read by every warp
        if (a.work) {
            if (threadIdx.x == 0) s_item = static_cast<int>(atomicAdd(a.work, 1u));
            __syncthreads();
            item = s_item;
        }
        if (item >= count) break;
        const uint32_t env = a.all_slots ? static_cast<uint32_t>(item % a.N)
                                         : (a.env_idx ? a.env_idx[item] : static_cast<uint32_t>(item));
        const size_t cells = static_cast<size_t>(G) * G;
        const uint32_t episode = a.all_slots ? a.episode_const + static_cast<uint32_t>(item / a.N)
                                             : (a.episode ? a.episode[item] : a.episode_const);
        const size_t plane = static_cast<size_t>(episode % a.S) * a.N + env;
        const uint32_t key = scenario_key(a.seed, a.env_id_base + env, episode);
        ScenarioParams sp;
        sp.si = sp.sj = sp.gi = sp.gj = 0; sp.yaw = 0.0f;
        if (lane == 0) {            // every warp samples the scenario for itself (no block barrier); warp 0 stores the record
            sp = sample_scenario(key, G, a.goal_mode);
            if (r == 0) store_scenario_record(a.scen_out + plane * SC_WORDS, sp, key);
        }
        sp.si = __shfl_sync(FULL, sp.si, 0); sp.sj = __shfl_sync(FULL, sp.sj, 0);
        sp.gi = __shfl_sync(FULL, sp.gi, 0); sp.gj = __shfl_sync(FULL, sp.gj, 0);
        const int gi = sp.gi, gj = sp.gj;

        uint32_t A[4], F[4], P0[4];
        scenario_free_row_il(key, lane * 4 + r, G, a.block_shift, a.p_thresh, sp, A);
        st4(&pl[pidx(IL_PFREE, r, lane)], A);
        {
            const uint32_t z[4] = {0u, 0u, 0u, 0u};
#pragma unroll
            for (int k = 1; k < IL_NPS; ++k) st4(&pl[pidx(k, r, lane)], z);
        }
        {
            const bool ok = gi >= 0 && gj >= 0 && gi < G && gj < G && lane == (gi >> 2) && r == (gi & 3);
            const int gw = gj & 3;
            const uint32_t bit = ok ? (1u << (gj >> 2)) : 0u;
#pragma unroll
            for (int w = 0; w < 4; ++w) {
                const uint32_t m = w == gw ? (bit & A[w]) : 0u;
                F[w] = m;
                A[w] ^= m;
                P0[w] = 0u;
            }
        }
        st4(&xch[0][slot], F);
        __syncthreads();

        // ---- the wavefront: one level per barrier ----
        int buf = 0;
        auto step = [&]() {
            uint32_t u[4], d[4], N[4];
            ld4(&xch[buf][up_slot], u);
            ld4(&xch[buf][dn_slot], d);
            const uint32_t lo0 = mask_on_fma(F[3], two), hi3 = F[0] >> 1;
            N[0] = (lo0 | F[1] | u[0] | d[0]) & A[0];
            N[1] = (F[0] | F[2] | u[1] | d[1]) & A[1];
            N[2] = (F[1] | F[3] | u[2] | d[2]) & A[2];
            N[3] = (F[2] | hi3 | u[3] | d[3]) & A[3];
            buf ^= 1;
            st4(&xch[buf][slot], N);
#pragma unroll
            for (int w = 0; w < 4; ++w) { A[w] = sub_on_fma(A[w], N[w], neg1); F[w] = N[w]; }
        };
        uint32_t L = 1;
        uint32_t s0 = one;
        for (;; L += 4) {
            // the plane schedule of the warp kernel: level L+1 toggles plane 0 (registers), level L+3 plane ctz(M), M = (L+3) >> 1
            const uint32_t M = (L + 3) >> 1;
            const int k = __ffs(M) - 1;
            const uint32_t sk = ((M >> (k + 1)) & 1u) ? neg1 : one;
            const bool resident = k < IL_NPS;
            uint32_t *pp = resident ? &pl[pidx(k, r, lane)] : &hi[pidx(k - IL_NPS, r, lane)];
            const bool first = M == (1u << k);                   // the scratch is not zeroed: its first toggle stores
            step();
            __syncthreads();
#pragma unroll
            for (int w = 0; w < 4; ++w) P0[w] = mad_u32(A[w], s0, P0[w]);
            s0 = 0u - s0;
            step();
            __syncthreads();
            uint32_t v[4] = {0u, 0u, 0u, 0u};
            if (resident || !first) ld4(pp, v);
            step();
            __syncthreads();
#pragma unroll
            for (int w = 0; w < 4; ++w) v[w] = mad_u32(A[w], sk, v[w]);
            st4(pp, v);
            step();
            if (!__syncthreads_or((F[0] | F[1] | F[2] | F[3]) != 0u)) break;
        }
        const uint32_t Mmax = (L + 2) >> 1;
        const int kmax = 32 - __clz(Mmax);
        {
            uint32_t v[4], f[4];
            ld4(&pl[pidx(IL_PFREE, r, lane)], f);
#pragma unroll
            for (int w = 0; w < 4; ++w) v[w] = f[w] & ~A[w];
            st4(&pl[pidx(IL_PVIS, r, lane)], v);
        }
        // ---- Gray -> binary of the thread's row, in place ----
        {
            uint32_t acc[4] = {0u, 0u, 0u, 0u};
#pragma unroll 1
            for (int k = kmax - 1; k >= IL_NPS; --k) {
                uint32_t v[4];
                uint32_t *p = &hi[pidx(k - IL_NPS, r, lane)];
                ld4(p, v);
#pragma unroll
                for (int w = 0; w < 4; ++w) acc[w] ^= v[w];
                st4(p, acc);
            }
#pragma unroll
            for (int k = IL_NPS - 1; k >= 1; --k) {
                uint32_t v[4];
                uint32_t *p = &pl[pidx(k, r, lane)];
                ld4(p, v);
#pragma unroll
                for (int w = 0; w < 4; ++w) acc[w] ^= v[w];
                st4(p, acc);
            }
#pragma unroll
            for (int w = 0; w < 4; ++w) acc[w] ^= P0[w];
            st4(&pl[pidx(0, r, lane)], acc);
        }
        __syncthreads();            // the output phase reads the rows above / below, which other warps converted
        il_emit_rows(pl, hi, stage[r], r, lane, G, gi, gj, kmax, a.flow + plane * cells, a.cost ? a.cost + plane * cells : nullptr);
    }

    // the last CTA to finish re-arms the regeneration list for its next use
    if (a.ticket) {
        __threadfence();
        __syncthreads();
        if (threadIdx.x == 0) flow_launch_epilogue(a);
    }
}

}  // namespace

size_t flow_field_il_scratch_words() { return static_cast<size_t>(IL_NPG) * IL_PW; }

// Resident warps per SM of a launch: 8 (2 per scheduler) measured best on 4096-grid batches (profiles/r02a_flow_ab.txt:
// 0.155 ms against 0.160 ms at 10 and 0.173 ms at 6) — shorter per-grid latency, hence a shorter tail of the launch.
int flow_field_il_ctas_per_sm() { return 8; }

// The four-warps-per-grid kernel serves the launches whose LATENCY is what the caller waits for: the flush of an unfinished
// regeneration group at a join (FlowArgs.latency) and resets of a few envs.  The regeneration launches that run BESIDE step
// kernels stay on the warp kernel: there the quad kernel's footprint (16 K registers per grid against 5.4 K) costs the step
// more than its shorter latency returns (profiles/r02e_quad_ab.txt: steady step 21.4 -> 23.4 us with every list on it).
// FFMP_FLOW_QUAD=0 keeps every launch on the warp kernel (the parity suite runs both); =k (1..4) sets the CTAs per SM.
static int quad_ctas_per_sm() {
    const char *e = std::getenv("FFMP_FLOW_QUAD");      // read per launch: the tests switch it between environments
    if (!e) return QUAD_CTAS_PER_SM;
    const int v = std::atoi(e);
    return v < 0 ? 0 : (v > 4 ? 4 : v);
}

cudaError_t launch_flow_field_il(const FlowArgs &a, int grid, cudaStream_t st) {
    const int qmax = 148 * quad_ctas_per_sm();
    if (a.generate && qmax > 0 && (a.latency || (!a.count_ptr && a.count <= qmax))) {
        flow_field_quad_kernel<<<grid < qmax ? grid : qmax, 128, 0, st>>>(a);
        return cudaGetLastError();
    }
    if (a.generate) flow_field_il_kernel<true><<<grid, 32, 0, st>>>(a);
    else flow_field_il_kernel<false><<<grid, 32, 0, st>>>(a);
    return cudaGetLastError();
}

}  // namespace ffmp
