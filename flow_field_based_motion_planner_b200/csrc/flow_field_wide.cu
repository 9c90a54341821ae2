// flow_field_wide.cu — SPEC.md §4/§5 for the large maps of BASELINE config 4: 384 < G <= 512 (G % 32 == 0).
// Replaces the external /bev/* flow-image ROS node (/root/reference/src/train.py:84,116-121).
//
// One CTA of FOUR WARPS per grid.  The layout is the one of flow_field_il.cu scaled up: a row of 512 cells is 16 words in the
// column-interleaved form (word w = the columns c % 16 == w, bit b <-> column 16 b + w), thread t of the CTA owns rows
// 4t .. 4t+3 as 64 + 64 registers (avail, frontier).  Per BFS level a thread does 2 LOP3 + 1 IMAD per word, the horizontal
// neighbours are the neighbouring word registers (2 shifts per row), the vertical ones the neighbouring row registers —
// 32 shuffles per warp for the lane-boundary rows — and only the three warp boundaries go through shared memory, ordered by
// ONE block barrier per level.  (Two rows per thread and eight warps — no register pressure, twice the shuffles per word —
// was measured as well: 3.8 ms against 2.2 ms per 512 grids.)  (The round-1 kernel, one thread per row, exchanged every row through
// shared memory: 48 LSU wavefronts per warp and level.)
// Levels are recorded as Gray-code planes of M = level >> 1 updated by addition (cost bit 0 is the checkerboard colour); planes
// 0..2 live in shared memory (96 KB: two CTAs per SM), higher planes, the free and the reached masks in a per-CTA L2 scratch.
// A thread-block cluster per grid was evaluated on paper and rejected: a level is ~1600 warp instructions for the CTA, i.e.
// ~600 cycles, while one cluster barrier + DSMEM halo round trip costs 380 + 215 cycles (B300_MICROARCH.md,
// "CGA; DSMEM"): splitting a grid over two SMs would lengthen every level.
// Algorithmic HBM bytes: 6 B/cell (1 occupancy read + 4 cost write + 1 flow write).
#include <mutex>
#include <type_traits>

#include "flow_bits.cuh"
#include "flow_rowops.cuh"

namespace ffmp {

namespace {

constexpr int WD_W = 16;                      // words per row
constexpr int WD_T = 128;                     // threads per CTA
constexpr int WD_R = 4;                       // rows per thread (512 padded rows); 64 + 64 registers of avail / frontier
constexpr int WD_NW = WD_T / 32;              // warps per CTA
constexpr int WD_NPS = 3;                     // shared-memory Gray planes (bits 0..2 of M)
constexpr int WD_NPG = 16;                    // global Gray planes (bits 3..18 of M: depth < 2^20)
constexpr int WD_PLANE = WD_R * 4 * WD_T * 4; // words of one plane: [r][q][thread][4]
constexpr int WD_XB = 2 * WD_NW * 2 * WD_W + WD_W;   // warp-boundary rows: [buffer][warp][top / bottom][16], then one all-zero row
constexpr int WD_SMEM = (WD_NPS * WD_PLANE + WD_XB) * 4 + 64;

struct WideInfo {
    int item, gi, gj;
    unsigned long long plane;
    uint32_t key;
    ScenarioParams sp;
};

__device__ __forceinline__ uint32_t madw(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t r;
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
    return r;
}

template <bool GEN>
__global__ void __launch_bounds__(WD_T, 2) flow_field_wide_kernel(FlowArgs a) {
    extern __shared__ __align__(16) uint32_t sm[];
    __shared__ WideInfo info;
    uint32_t *const xb = sm + WD_NPS * WD_PLANE;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int G = a.G;
    const int count = a.count_ptr ? static_cast<int>(*a.count_ptr) : a.count;
    // per-CTA scratch (L2): Gray planes 3.., then the free and the reached masks
    uint32_t *const hi = a.hi_scratch + static_cast<size_t>(blockIdx.x) * ((WD_NPG + 2) * WD_PLANE);
    uint32_t *const free_g = hi + WD_NPG * WD_PLANE, *const vis_g = free_g + WD_PLANE;
    const size_t cells = static_cast<size_t>(G) * G;
    const uint32_t neg1 = a.neg1, one = a.one, two = a.one + a.one;
    // word (r, w) of thread t: 16-byte chunk q = w / 4 at ((r * 4 + q) * 128 + t) * 4: consecutive threads, consecutive chunks
    auto pidx = [](int r, int q, int t) { return ((r * 4 + q) * WD_T + t) * 4; };      // r < WD_R
    auto ldrow = [&](const uint32_t *pl, int r, int t, uint32_t (&v)[WD_W]) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const uint4 x = *reinterpret_cast<const uint4 *>(pl + pidx(r, q, t));
            v[4 * q] = x.x; v[4 * q + 1] = x.y; v[4 * q + 2] = x.z; v[4 * q + 3] = x.w;
        }
    };
    auto strow = [&](uint32_t *pl, int r, int t, const uint32_t (&v)[WD_W]) {
#pragma unroll
        for (int q = 0; q < 4; ++q)
            *reinterpret_cast<uint4 *>(pl + pidx(r, q, t)) = make_uint4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
    };
    auto gray_plane = [&](int k) -> uint32_t * { return k < WD_NPS ? sm + k * WD_PLANE : hi + (k - WD_NPS) * WD_PLANE; };

    for (int item = blockIdx.x;; item += gridDim.x) {
        __syncthreads();                      // the previous grid's use of `info` and of the planes is over
        if (a.work) {
            if (tid == 0) info.item = static_cast<int>(atomicAdd(a.work, 1u));
            __syncthreads();
            item = info.item;
        }
        if (item >= count) break;
        if (tid == 0) {
            const uint32_t env = a.env_idx ? a.env_idx[item] : static_cast<uint32_t>(item);
            if (GEN) {
                const uint32_t episode = a.episode ? a.episode[item] : a.episode_const;
                info.plane = static_cast<unsigned long long>(episode % a.S) * a.N + env;
                info.key = scenario_key(a.seed, a.env_id_base + env, episode);
                info.sp = sample_scenario(info.key, G, a.goal_mode);
                store_scenario_record(a.scen_out + info.plane * SC_WORDS, info.sp, info.key);
                info.gi = info.sp.gi; info.gj = info.sp.gj;
            } else if (a.slot_mode) {
                info.plane = static_cast<unsigned long long>((a.episode ? a.episode[item] : a.episode_const) % a.S) * a.N + env;
                info.gi = static_cast<int>(a.scen[info.plane * SC_WORDS + SC_GI]);
                info.gj = static_cast<int>(a.scen[info.plane * SC_WORDS + SC_GJ]);
            } else {
                info.plane = static_cast<unsigned long long>(item);
                info.gi = a.goal_cells[2 * item];
                info.gj = a.goal_cells[2 * item + 1];
            }
        }
        __syncthreads();
        const size_t plane = static_cast<size_t>(info.plane);
        const int gi = info.gi, gj = info.gj;

        // ---- 1. free mask of the thread's four rows (registers + the L2 scratch), zeroed shared-memory planes ----
        uint32_t A[WD_R][WD_W], F[WD_R][WD_W];
#pragma unroll 1
        for (int r = 0; r < WD_R; ++r) {
            const int R = WD_R * tid + r;
            uint32_t fr[WD_W];
            if (GEN) {
                // generated scenarios (SPEC.md §3): the linear 32-column words of the row, then linear -> interleaved as a
                // 32 x 16 bit transpose of the row's 32 halfwords (halfword q = columns 16 q .. 16 q + 15)
                uint32_t x[16];
#pragma unroll 1
                for (int q = 0; q < 16; ++q) {
                    const uint32_t lo = scenario_free_word(info.key, R, 32 * (q >> 1), G, a.block_shift, a.p_thresh, info.sp);
                    const uint32_t hw = scenario_free_word(info.key, R, 32 * ((q + 16) >> 1), G, a.block_shift, a.p_thresh, info.sp);
                    const uint32_t v = ((q & 1) ? lo >> 16 : lo & 0xFFFFu) | (((q & 1) ? hw >> 16 : hw & 0xFFFFu) << 16);
#pragma unroll
                    for (int u = 0; u < 16; ++u)
                        if (u == q) x[u] = v;
                }
                rowops::transpose16x2(x);
#pragma unroll
                for (int w = 0; w < WD_W; ++w) fr[w] = x[w];
            } else if (R < G) {
                const uint8_t *src = a.occ + plane * cells + static_cast<size_t>(R) * G;
                if (G == 512) {
                    // 16 bytes = the columns 16 q .. 16 q + 15 = bit q of the 16 words: the row is the bit transpose of its 32
                    // 16-bit occupancy flags
                    uint32_t x[16];
#pragma unroll
                    for (int q = 0; q < 16; ++q) {
                        const uint4 lo = __ldg(reinterpret_cast<const uint4 *>(src) + q), hw = __ldg(reinterpret_cast<const uint4 *>(src) + q + 16);
                        x[q] = rowops::occupied_flags16(lo.x, lo.y, lo.z, lo.w) | (rowops::occupied_flags16(hw.x, hw.y, hw.z, hw.w) << 16);
                    }
                    rowops::transpose16x2(x);
#pragma unroll
                    for (int w = 0; w < WD_W; ++w) fr[w] = ~x[w];
                } else {
                    // 384 < G < 512: word by word (static register index), a byte per bit
#pragma unroll
                    for (int w = 0; w < WD_W; ++w) {
                        uint32_t v = 0;
                        for (int b = 0; b * WD_W + w < G; ++b) v |= (__ldg(src + b * WD_W + w) == 0 ? 1u : 0u) << b;
                        fr[w] = v;
                    }
                }
            } else {
#pragma unroll
                for (int w = 0; w < WD_W; ++w) fr[w] = 0u;
            }
            strow(free_g, r, tid, fr);
            const uint32_t z[WD_W] = {0u};
#pragma unroll
            for (int k = 0; k < WD_NPS; ++k) strow(sm + k * WD_PLANE, r, tid, z);
#pragma unroll
            for (int u = 0; u < WD_R; ++u)
                if (u == r) {
#pragma unroll
                    for (int w = 0; w < WD_W; ++w) A[u][w] = fr[w];
                }
        }
        for (int i = tid; i < WD_XB; i += WD_T) xb[i] = 0u;
        {
            const bool ok = gi >= 0 && gj >= 0 && gi < G && gj < G && tid == gi / WD_R;
            const int gr = gi % WD_R, gw = gj & 15;
            const uint32_t bit = ok ? (1u << (gj >> 4)) : 0u;
#pragma unroll
            for (int r = 0; r < WD_R; ++r)
#pragma unroll
                for (int w = 0; w < WD_W; ++w) {
                    const uint32_t m = (r == gr && w == gw) ? (bit & A[r][w]) : 0u;
                    F[r][w] = m;
                    A[r][w] ^= m;
                }
        }
        __syncthreads();

        // ---- 2. bit-parallel wavefront ----
        int cur = 0;
        auto xrow = [&](int buf, int w_, int side) -> uint32_t * { return xb + ((buf * WD_NW + w_) * 2 + side) * WD_W; };
        const uint32_t *const zero_row = xb + 2 * WD_NW * 2 * WD_W;
        // the warp's first / last rows for the neighbouring warps (read after the level's barrier)
        auto publish = [&](const uint32_t (&top)[WD_W], const uint32_t (&bottom)[WD_W], int buf) {
            if (lane == 0) {
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    *reinterpret_cast<uint4 *>(xrow(buf, warp, 0) + 4 * q) = make_uint4(top[4 * q], top[4 * q + 1], top[4 * q + 2], top[4 * q + 3]);
            }
            if (lane == 31) {
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    *reinterpret_cast<uint4 *>(xrow(buf, warp, 1) + 4 * q) =
                        make_uint4(bottom[4 * q], bottom[4 * q + 1], bottom[4 * q + 2], bottom[4 * q + 3]);
            }
        };
        auto lo = [&](const uint32_t (&x)[WD_W], int w) { return w > 0 ? x[w - 1] : mask_on_fma(x[WD_W - 1], two); };
        auto hh = [&](const uint32_t (&x)[WD_W], int w) { return w < WD_W - 1 ? x[w + 1] : x[0] >> 1; };
        // One BFS level.  Register budget (255 per thread at two CTAs per SM): avail 64 + frontier 64 + two 16-word temporaries
        // + the two neighbour rows = 208; the rows are therefore renewed in the order N0, N1 | F0 <- N0 | N2 | F1 <- N1 | N3 |
        // F2 <- N2, F3 <- N3, which needs only two temporaries while every row still sees its neighbours' OLD frontier.
        auto step = [&]() {
            uint32_t u[WD_W], d[WD_W];
            {
                // rows 4t - 1 / 4t + 4: the neighbouring lanes' (shuffles), or for the warp's edge lanes the row the neighbouring
                // warp published before the previous barrier (every lane reads it: a broadcast, no divergent branch; the
                // grid's first / last warp reads the all-zero row)
                const uint32_t *pu = warp > 0 ? xrow(cur, warp - 1, 1) : zero_row;
                const uint32_t *pd = warp < WD_NW - 1 ? xrow(cur, warp + 1, 0) : zero_row;
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const uint4 tu4 = *reinterpret_cast<const uint4 *>(pu + 4 * q), td4 = *reinterpret_cast<const uint4 *>(pd + 4 * q);
                    const uint32_t xu[4] = {tu4.x, tu4.y, tu4.z, tu4.w}, xd[4] = {td4.x, td4.y, td4.z, td4.w};
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const uint32_t su = __shfl_up_sync(FULL, F[3][4 * q + j], 1), sd = __shfl_down_sync(FULL, F[0][4 * q + j], 1);
                        u[4 * q + j] = lane == 0 ? xu[j] : su;
                        d[4 * q + j] = lane == 31 ? xd[j] : sd;
                    }
                }
            }
            // two passes per row (all first halves, then all second halves): with one or two warps per scheduler a dependent
            // LOP3 pair issued back to back would wait out the 4-cycle ALU latency 64 times per level
            uint32_t T0[WD_W], T1[WD_W];
#pragma unroll
            for (int w = 0; w < WD_W; ++w) T0[w] = lo(F[0], w) | hh(F[0], w) | u[w];
#pragma unroll
            for (int w = 0; w < WD_W; ++w) T1[w] = lo(F[1], w) | hh(F[1], w) | F[0][w];
#pragma unroll
            for (int w = 0; w < WD_W; ++w) T0[w] = (T0[w] | F[1][w]) & A[0][w];
#pragma unroll
            for (int w = 0; w < WD_W; ++w) T1[w] = (T1[w] | F[2][w]) & A[1][w];
#pragma unroll
            for (int w = 0; w < WD_W; ++w) { A[0][w] = sub_on_fma(A[0][w], T0[w], neg1); F[0][w] = T0[w]; }
#pragma unroll
            for (int w = 0; w < WD_W; ++w) T0[w] = lo(F[2], w) | hh(F[2], w) | F[1][w];
#pragma unroll
            for (int w = 0; w < WD_W; ++w) { A[1][w] = sub_on_fma(A[1][w], T1[w], neg1); F[1][w] = T1[w]; }
#pragma unroll
            for (int w = 0; w < WD_W; ++w) T1[w] = lo(F[3], w) | hh(F[3], w) | d[w];
#pragma unroll
            for (int w = 0; w < WD_W; ++w) T0[w] = (T0[w] | F[3][w]) & A[2][w];
#pragma unroll
            for (int w = 0; w < WD_W; ++w) T1[w] = (T1[w] | F[2][w]) & A[3][w];
#pragma unroll
            for (int w = 0; w < WD_W; ++w) {
                A[2][w] = sub_on_fma(A[2][w], T0[w], neg1); F[2][w] = T0[w];
                A[3][w] = sub_on_fma(A[3][w], T1[w], neg1); F[3][w] = T1[w];
            }
            publish(F[0], F[3], cur ^ 1);
            cur ^= 1;
            __syncthreads();
        };
        // level L even, before its step: plane ctz(L >> 1) += +-avail (nested sets: the alternating sum is the XOR)
        auto gray = [&](uint32_t L) {
            const uint32_t M = L >> 1;
            const int k = __ffs(M) - 1;
            const uint32_t s = ((M >> (k + 1)) & 1u) ? neg1 : one;
            if (k < WD_NPS) {
                uint32_t *pl = sm + k * WD_PLANE;
#pragma unroll
                for (int r = 0; r < WD_R; ++r) {
                    uint32_t v[WD_W];
                    ldrow(pl, r, tid, v);
#pragma unroll
                    for (int w = 0; w < WD_W; ++w) v[w] = madw(A[r][w], s, v[w]);
                    strow(pl, r, tid, v);
                }
            } else {
                uint32_t *pl = hi + (k - WD_NPS) * WD_PLANE;
                const bool first = M == (1u << k);      // the scratch is not zeroed: its first toggle stores
#pragma unroll
                for (int r = 0; r < WD_R; ++r) {
                    uint32_t v[WD_W];
                    if (first) {
#pragma unroll
                        for (int w = 0; w < WD_W; ++w) v[w] = 0u;
                    } else {
                        ldrow(pl, r, tid, v);
                    }
#pragma unroll
                    for (int w = 0; w < WD_W; ++w) v[w] = madw(A[r][w], s, v[w]);
                    strow(pl, r, tid, v);
                }
            }
        };
        publish(F[0], F[WD_R - 1], cur);
        __syncthreads();
        uint32_t L = 1;
        for (;; L += 4) {
#pragma unroll 1
            for (uint32_t h = 0; h < 4; h += 2) {
                step();
                gray(L + h + 1);
                step();
            }
            uint32_t any = 0;
#pragma unroll
            for (int r = 0; r < WD_R; ++r)
#pragma unroll
                for (int w = 0; w < WD_W; ++w) any |= F[r][w];
            if (!__syncthreads_or(any != 0)) break;
        }
        const uint32_t Mmax = (L + 2) >> 1;
        const int kmax = 32 - __clz(Mmax);

        // ---- 3. Gray -> binary in place (plane k becomes cost bit k + 1); the reached mask to the scratch ----
#pragma unroll 1
        for (int r = 0; r < WD_R; ++r) {
            uint32_t acc[WD_W] = {0u}, fr[WD_W], v[WD_W];
#pragma unroll 1
            for (int k = kmax - 1; k >= 0; --k) {
                uint32_t *pl = gray_plane(k);
                ldrow(pl, r, tid, v);
#pragma unroll
                for (int w = 0; w < WD_W; ++w) acc[w] ^= v[w];
                strow(pl, r, tid, acc);
            }
            if (kmax < WD_NPS) {                // planes the flow direction reads must be defined
                const uint32_t z[WD_W] = {0u};
                for (int k = kmax; k < 2; ++k) strow(sm + k * WD_PLANE, r, tid, z);
            }
            ldrow(free_g, r, tid, fr);
#pragma unroll
            for (int u = 0; u < WD_R; ++u)
                if (u == r) {
#pragma unroll
                    for (int w = 0; w < WD_W; ++w) v[w] = fr[w] & ~A[u][w];
                }
            strow(vis_g, r, tid, v);
        }
        __syncthreads();

        // ---- 4. per row: flow direction -> flow bytes, cost planes -> int32 (word group by word group) ----
        uint8_t *flow = a.flow + plane * cells;
        int32_t *cost = a.cost ? a.cost + plane * cells : nullptr;
        const uint32_t *b1p = sm, *b2p = sm + WD_PLANE;
#pragma unroll 1
        for (int r = 0; r < WD_R; ++r) {
            const int R = WD_R * tid + r;
            const int tu = r == 0 ? tid - 1 : tid, ru = r == 0 ? WD_R - 1 : r - 1;
            const int td = r == WD_R - 1 ? tid + 1 : tid, rd = r == WD_R - 1 ? 0 : r + 1;
            const uint32_t par0 = static_cast<uint32_t>(R + gi + gj) & 1u;
            rowops::RowInW<WD_W> in;
            ldrow(b1p, r, tid, in.b1c); ldrow(b2p, r, tid, in.b2c); ldrow(vis_g, r, tid, in.Vc); ldrow(free_g, r, tid, in.Fc);
#pragma unroll
            for (int w = 0; w < WD_W; ++w) { in.b1u[w] = in.b2u[w] = in.Vu[w] = 0u; in.b1d[w] = in.b2d[w] = in.Vd[w] = 0u; }
            if (tu >= 0) { ldrow(b1p, ru, tu, in.b1u); ldrow(b2p, ru, tu, in.b2u); ldrow(vis_g, ru, tu, in.Vu); }
            if (td < WD_T) { ldrow(b1p, rd, td, in.b1d); ldrow(b2p, rd, td, in.b2d); ldrow(vis_g, rd, td, in.Vd); }
            uint32_t n[4][WD_W];
            rowops::direction_nibbles_w<WD_W>(in, par0, n);
            if (R >= G) continue;
#pragma unroll 1
            for (int j = 0; j < 4; ++j) {
                uint32_t nj[4][4];
#pragma unroll
                for (int u = 0; u < 4; ++u)
                    if (u == j) {
#pragma unroll
                        for (int q = 0; q < 4; ++q)
#pragma unroll
                            for (int wl = 0; wl < 4; ++wl) nj[q][wl] = n[q][4 * u + wl];
                    }
                uint32_t fw[32];
                rowops::flow_row_words(nj, fw);           // word b = the 4 cells at columns 16 b + 4 j
                uint8_t *fdst = flow + static_cast<size_t>(R) * G + 4 * j;
#pragma unroll
                for (int b = 0; b < 32; ++b)
                    if (16 * b + 4 * j < G) *reinterpret_cast<uint32_t *>(fdst + 16 * b) = fw[b];
                if (!cost) continue;
                int32_t *cdst = cost + static_cast<size_t>(R) * G + 4 * j;
                uint32_t Vj[4];
#pragma unroll
                for (int u = 0; u < 4; ++u)
                    if (u == j) {
#pragma unroll
                        for (int wl = 0; wl < 4; ++wl) Vj[wl] = in.Vc[4 * u + wl];
                    }
                const uint32_t par = (par0 & 1u) ? 0x00010001u : 0x01000100u;      // 4 j is even: the group starts on par0's colour
                if (kmax <= 15) {
                    uint32_t xl[32], xh[32];
#pragma unroll
                    for (int wl = 0; wl < 4; ++wl) xl[8 * wl] = ~Vj[wl];
#pragma unroll
                    for (int k = 0; k < 7; ++k) {       // planes >= kmax were never written (the L2 scratch is not zeroed)
                        uint4 t = make_uint4(0u, 0u, 0u, 0u);
                        if (k < kmax) t = *reinterpret_cast<const uint4 *>(gray_plane(k) + pidx(r, j, tid));
                        xl[1 + k] = t.x | ~Vj[0]; xl[9 + k] = t.y | ~Vj[1]; xl[17 + k] = t.z | ~Vj[2]; xl[25 + k] = t.w | ~Vj[3];
                    }
                    rowops::transpose32(xl);
                    if (kmax > 7) {
#pragma unroll
                        for (int k = 0; k < 8; ++k) {
                            uint4 t = make_uint4(0u, 0u, 0u, 0u);
                            if (7 + k < kmax) t = *reinterpret_cast<const uint4 *>(gray_plane(7 + k) + pidx(r, j, tid));
                            xh[k] = t.x; xh[8 + k] = t.y; xh[16 + k] = t.z; xh[24 + k] = t.w;
                        }
                        rowops::transpose32(xh);
                    } else {
#pragma unroll
                        for (int b = 0; b < 32; ++b) xh[b] = 0u;
                    }
#pragma unroll
                    for (int b = 0; b < 32; ++b)
                        if (16 * b + 4 * j < G) {
                            const rowops::Int4 c = rowops::widen_cost4_16(xl[b], xh[b], par);
                            *reinterpret_cast<int4 *>(cdst + 16 * b) = make_int4(c.x, c.y, c.z, c.w);
                        }
                } else {
                    // mazes deeper than 2^16 levels: cell by cell from the planes
#pragma unroll 1
                    for (int b = 0; 16 * b + 4 * j < G; ++b) {
                        uint32_t v[4];
#pragma unroll
                        for (int wl = 0; wl < 4; ++wl) v[wl] = (par0 ^ static_cast<uint32_t>(wl)) & 1u;
#pragma unroll 1
                        for (int k = 0; k < kmax; ++k) {
                            const uint4 t = *reinterpret_cast<const uint4 *>(gray_plane(k) + pidx(r, j, tid));
                            v[0] |= ((t.x >> b) & 1u) << (k + 1); v[1] |= ((t.y >> b) & 1u) << (k + 1);
                            v[2] |= ((t.z >> b) & 1u) << (k + 1); v[3] |= ((t.w >> b) & 1u) << (k + 1);
                        }
                        int4 c;
                        c.x = (Vj[0] >> b) & 1u ? static_cast<int>(v[0]) : COST_INF;
                        c.y = (Vj[1] >> b) & 1u ? static_cast<int>(v[1]) : COST_INF;
                        c.z = (Vj[2] >> b) & 1u ? static_cast<int>(v[2]) : COST_INF;
                        c.w = (Vj[3] >> b) & 1u ? static_cast<int>(v[3]) : COST_INF;
                        *reinterpret_cast<int4 *>(cdst + 16 * b) = c;
                    }
                }
            }
        }
    }

    // the last CTA to finish re-arms the regeneration list for its next use
    if (a.ticket && tid == 0) {
        __threadfence();
        const uint32_t t = atomicAdd(a.ticket, 1u);
        if (t == gridDim.x - 1) {
            *a.ticket = 0;
            if (a.work) *a.work = 0;
            if (a.count_reset) *a.count_reset = 0;
            __threadfence();
            if (a.host_done) {
                __threadfence_system();
                *reinterpret_cast<volatile uint32_t *>(a.host_done) = a.host_done_value;
            }
        }
    }
}

}  // namespace

bool flow_field_wide_supported(int G) { return G > 384 && G <= 512 && (G % 32) == 0; }
size_t flow_field_wide_scratch_words() { return static_cast<size_t>(WD_NPG + 2) * WD_PLANE; }
int flow_field_wide_max_grid() { return 148 * 2; }

cudaError_t launch_flow_field_wide(const FlowArgs &a_in, int grid, cudaStream_t st) {
    if (grid <= 0) return cudaSuccess;
    FlowArgs a = a_in;
    a.neg1 = 0xFFFFFFFFu;
    a.one = 1u;
    // the opt-in for > 48 KB of dynamic shared memory belongs to the device of the call: tracked per ordinal
    static std::mutex mu;
    static bool configured_dev[64] = {false};
    int dev = 0;
    if (cudaError_t ce = cudaGetDevice(&dev); ce != cudaSuccess) return ce;
    {
        std::lock_guard<std::mutex> lock(mu);
        const bool known = dev >= 0 && dev < 64;
        if (!known || !configured_dev[dev]) {
            cudaError_t ce = cudaFuncSetAttribute(flow_field_wide_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, WD_SMEM);
            if (ce == cudaSuccess) ce = cudaFuncSetAttribute(flow_field_wide_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, WD_SMEM);
            if (ce != cudaSuccess) return ce;
            if (known) configured_dev[dev] = true;
        }
    }
    if (a.generate) flow_field_wide_kernel<true><<<grid, WD_T, WD_SMEM, st>>>(a);
    else flow_field_wide_kernel<false><<<grid, WD_T, WD_SMEM, st>>>(a);
    return cudaGetLastError();
}

}  // namespace ffmp
