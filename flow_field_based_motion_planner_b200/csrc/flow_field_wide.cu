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
// A thread-block cluster per grid (two CTAs, rows split at 256) exists as the CL = 2 instantiation of the same kernel
// (FFMP_FLOW_CLUSTER=1): the seam row travels by st.async into the peer's shared memory and completes transaction bytes on the
// peer's mbarrier, the convergence vote and the grid hand-out are cluster-wide.  It is bit-exact and slower (1.35 ms against
// 0.74 ms for 16 grids, 3.0 against 2.2 ms for 512; profiles/r02f_cluster_ab.txt), so it is not the default — DESIGN.md §3.3.
// Algorithmic HBM bytes: 6 B/cell (1 occupancy read + 4 cost write + 1 flow write).
#include <cstdlib>
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
constexpr int WD_XR = 2 * WD_W;               // cluster variant: [buffer][16], the seam row the peer CTA sends (st.async into this CTA's shared memory)
constexpr int WD_SMEM = (WD_NPS * WD_PLANE + WD_XB) * 4 + 64;
constexpr int WD_SMEM_CL2 = (WD_NPS * (WD_PLANE / 2) + WD_XB + WD_XR) * 4 + 64;

// ---- thread-block cluster helpers (CL = 2: one grid per CTA PAIR, rows 0..255 on rank 0, 256..511 on rank 1) ----
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
// barrier over every thread of the cluster; release / acquire at cluster scope (shared memory of both CTAs and global memory)
__device__ __forceinline__ void cluster_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// 32-bit shared::cluster address of `p` (a shared-memory object of this CTA) in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t peer_shared_u32(const void *p, uint32_t rank) {
    uint32_t out;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(out) : "r"(static_cast<uint32_t>(__cvta_generic_to_shared(p))), "r"(rank));
    return out;
}
// 16 bytes into the peer's shared memory, counted as 16 transaction bytes on the peer's mbarrier (no cluster-wide barrier, no fence)
__device__ __forceinline__ void st_async16(uint32_t dst, uint32_t x, uint32_t y, uint32_t z, uint32_t w, uint32_t mbar) {
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];"
                 ::"r"(dst), "r"(x), "r"(y), "r"(z), "r"(w), "r"(mbar) : "memory");
}
__device__ __forceinline__ void st_async4(uint32_t dst, uint32_t x, uint32_t mbar) {
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b32 [%0], %1, [%2];" ::"r"(dst), "r"(x), "r"(mbar) : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar, uint32_t parity) {
    // a non-blocking test in a spin loop: try_wait may park the warp for a system-defined time slice
    uint32_t done;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    } while (!done);
}
// generic address of `p` (a shared-memory address of this CTA) in the shared memory of CTA `rank` of the cluster
template <typename T>
__device__ __forceinline__ T *peer_shared(T *p, uint32_t rank) {
    uint64_t out;
    asm volatile("mapa.u64 %0, %1, %2;" : "=l"(out) : "l"(reinterpret_cast<uint64_t>(p)), "r"(rank));
    return reinterpret_cast<T *>(out);
}

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

// CL = 1: one CTA per grid (the default).  CL = 2: a thread-block cluster of two CTAs per grid — the "multi-CTA wavefront per
// env" of BASELINE config 4: CTA rank c owns rows 256c .. 256c + 255 (thread t: rows 2 (128c + t), + 1), the row at the seam
// travels into the peer's shared memory with a DSMEM store per level, `barrier.cluster` replaces the block barrier, the
// convergence vote and the grid hand-out are cluster-wide, and the output phase reads the seam's neighbour rows from the
// peer's planes.  Measured against CL = 1 in profiles/r02f_cluster_ab.txt.
template <bool GEN, int CL>
__global__ void __launch_bounds__(WD_T, 2) flow_field_wide_kernel(FlowArgs a) {
    constexpr int RPT = WD_R / CL;                    // rows per thread
    constexpr int PLANE = WD_PLANE / CL;              // words of one plane held by this CTA
    extern __shared__ __align__(16) uint32_t sm[];
    __shared__ WideInfo info;
    __shared__ uint32_t conv_in[2];                   // cluster variant: [vote parity] the peer's "my half still has a frontier"
    __shared__ __align__(8) uint64_t mb_row[2], mb_flag[2];   // ... and the mbarriers that count the peer's bytes: seam rows, votes
    uint32_t *const xb = sm + WD_NPS * PLANE;
    uint32_t *const xr = xb + WD_XB;                  // cluster variant: the seam rows the peer publishes
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t crank = CL > 1 ? cluster_ctarank() : 0u;
    const int gtid = static_cast<int>(crank) * WD_T + tid;          // thread index within the grid's CL * 128 threads
    const int G = a.G;
    const int count = a.count_ptr ? static_cast<int>(*a.count_ptr) : a.count;
    // per-CTA scratch (L2): Gray planes 3.., then the free and the reached masks
    uint32_t *const hi = a.hi_scratch + static_cast<size_t>(blockIdx.x) * ((WD_NPG + 2) * WD_PLANE);
    uint32_t *const free_g = hi + WD_NPG * WD_PLANE, *const vis_g = free_g + WD_PLANE;
    const uint32_t *const vis_peer = a.hi_scratch + static_cast<size_t>(blockIdx.x ^ 1u) * ((WD_NPG + 2) * WD_PLANE) + (WD_NPG + 1) * WD_PLANE;
    const uint32_t *const sm_peer = CL > 1 ? peer_shared(sm, crank ^ 1u) : sm;
    // the seam: rank 0's last row is "the row above" of rank 1, rank 1's first row "the row below" of rank 0.  Each level the
    // owner sends its row with four st.async (16 bytes each, completing transaction bytes on the RECEIVER's mbarrier); only the
    // receiver's seam warp waits for it.  The rows are double-buffered by level parity; a sender cannot overwrite a buffer the
    // receiver still reads, because its next-but-one row depends on the receiver's answer to the row in between.
    const bool seam_warp = CL > 1 && ((crank == 1 && warp == 0) || (crank == 0 && warp == WD_NW - 1));
    const bool seam_lane = seam_warp && ((crank == 1 && lane == 0) || (crank == 0 && lane == 31));
    const uint32_t xr_peer32 = CL > 1 ? peer_shared_u32(xr, crank ^ 1u) : 0u;
    const uint32_t mb_row_peer32 = CL > 1 ? peer_shared_u32(mb_row, crank ^ 1u) : 0u;
    const uint32_t mb_row32 = static_cast<uint32_t>(__cvta_generic_to_shared(mb_row));
    const uint32_t mb_flag32 = static_cast<uint32_t>(__cvta_generic_to_shared(mb_flag));
    uint32_t ph_row0 = 0, ph_row1 = 0, ph_flag0 = 0, ph_flag1 = 0;      // phase parities of the four mbarriers
    if constexpr (CL > 1) {
        if (tid == 0) {
            for (int i = 0; i < 2; ++i) { mbar_init(mb_row32 + 8 * i, 1); mbar_init(mb_flag32 + 8 * i, 1); }
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
    }
    auto seam_wait = [&](int buf) {       // seam warp: arm the buffer's mbarrier for the peer's 64 bytes and wait for them
        if (lane == 0) mbar_expect_tx(mb_row32 + 8 * buf, 4 * WD_W);
        const uint32_t ph = buf ? ph_row1 : ph_row0;
        mbar_wait_cluster(mb_row32 + 8 * buf, ph);
        if (buf) ph_row1 ^= 1u; else ph_row0 ^= 1u;
    };
    const size_t cells = static_cast<size_t>(G) * G;
    const uint32_t neg1 = a.neg1, one = a.one, two = a.one + a.one;
    auto block_or_cluster_sync = [&]() {
        if constexpr (CL > 1) cluster_sync();
        else __syncthreads();
    };
    // word (r, w) of thread t: 16-byte chunk q = w / 4 at ((r * 4 + q) * 128 + t) * 4: consecutive threads, consecutive chunks
    auto pidx = [](int r, int q, int t) { return ((r * 4 + q) * WD_T + t) * 4; };      // r < RPT
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
    auto gray_plane = [&](int k) -> uint32_t * { return k < WD_NPS ? sm + k * PLANE : hi + (k - WD_NPS) * WD_PLANE; };

    for (int item = blockIdx.x / CL;; item += gridDim.x / CL) {
        block_or_cluster_sync();              // the previous grid's use of `info` and of the planes (the peer's reads too) is over
        if (a.work) {
            if (tid == 0 && crank == 0) {
                const int it = static_cast<int>(atomicAdd(a.work, 1u));
                info.item = it;
                if constexpr (CL > 1) peer_shared(&info, 1u)->item = it;
            }
            block_or_cluster_sync();
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
                if (crank == 0) store_scenario_record(a.scen_out + info.plane * SC_WORDS, info.sp, info.key);
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
        uint32_t A[RPT][WD_W], F[RPT][WD_W];
#pragma unroll 1
        for (int r = 0; r < RPT; ++r) {
            const int R = RPT * gtid + r;
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
            for (int k = 0; k < WD_NPS; ++k) strow(sm + k * PLANE, r, tid, z);
#pragma unroll
            for (int u = 0; u < RPT; ++u)
                if (u == r) {
#pragma unroll
                    for (int w = 0; w < WD_W; ++w) A[u][w] = fr[w];
                }
        }
        for (int i = tid; i < WD_XB; i += WD_T) xb[i] = 0u;          // (the seam rows `xr` are always overwritten whole by the peer)
        {
            const bool ok = gi >= 0 && gj >= 0 && gi < G && gj < G && gtid == gi / RPT;
            const int gr = gi % RPT, gw = gj & 15;
            const uint32_t bit = ok ? (1u << (gj >> 4)) : 0u;
#pragma unroll
            for (int r = 0; r < RPT; ++r)
#pragma unroll
                for (int w = 0; w < WD_W; ++w) {
                    const uint32_t m = (r == gr && w == gw) ? (bit & A[r][w]) : 0u;
                    F[r][w] = m;
                    A[r][w] ^= m;
                }
        }
        block_or_cluster_sync();              // the exchange rows are zeroed before anybody (the peer included) publishes into them

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
            if constexpr (CL > 1) {
                if (seam_lane) {
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        if (crank == 0) st_async16(xr_peer32 + (buf * WD_W + 4 * q) * 4, bottom[4 * q], bottom[4 * q + 1], bottom[4 * q + 2], bottom[4 * q + 3], mb_row_peer32 + 8 * buf);
                        else st_async16(xr_peer32 + (buf * WD_W + 4 * q) * 4, top[4 * q], top[4 * q + 1], top[4 * q + 2], top[4 * q + 3], mb_row_peer32 + 8 * buf);
                    }
                }
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
                if constexpr (CL > 1) {
                    if (seam_warp) seam_wait(cur);
                }
                const uint32_t *pu = warp > 0 ? xrow(cur, warp - 1, 1) : ((CL > 1 && crank == 1) ? xr + cur * WD_W : zero_row);
                const uint32_t *pd = warp < WD_NW - 1 ? xrow(cur, warp + 1, 0) : ((CL > 1 && crank == 0) ? xr + cur * WD_W : zero_row);
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const uint4 tu4 = *reinterpret_cast<const uint4 *>(pu + 4 * q), td4 = *reinterpret_cast<const uint4 *>(pd + 4 * q);
                    const uint32_t xu[4] = {tu4.x, tu4.y, tu4.z, tu4.w}, xd[4] = {td4.x, td4.y, td4.z, td4.w};
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const uint32_t su = __shfl_up_sync(FULL, F[RPT - 1][4 * q + j], 1), sd = __shfl_down_sync(FULL, F[0][4 * q + j], 1);
                        u[4 * q + j] = lane == 0 ? xu[j] : su;
                        d[4 * q + j] = lane == 31 ? xd[j] : sd;
                    }
                }
            }
            // two passes per row (all first halves, then all second halves): with one or two warps per scheduler a dependent
            // LOP3 pair issued back to back would wait out the 4-cycle ALU latency 64 times per level
            uint32_t T0[WD_W], T1[WD_W];
            if constexpr (RPT == 4) {
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
            } else {
                // two rows per thread (cluster variant): both rows are lane-boundary rows
#pragma unroll
                for (int w = 0; w < WD_W; ++w) T0[w] = lo(F[0], w) | hh(F[0], w) | u[w];
#pragma unroll
                for (int w = 0; w < WD_W; ++w) T1[w] = lo(F[1], w) | hh(F[1], w) | d[w];
#pragma unroll
                for (int w = 0; w < WD_W; ++w) T0[w] = (T0[w] | F[1][w]) & A[0][w];
#pragma unroll
                for (int w = 0; w < WD_W; ++w) T1[w] = (T1[w] | F[0][w]) & A[1][w];
#pragma unroll
                for (int w = 0; w < WD_W; ++w) {
                    A[0][w] = sub_on_fma(A[0][w], T0[w], neg1); F[0][w] = T0[w];
                    A[1][w] = sub_on_fma(A[1][w], T1[w], neg1); F[1][w] = T1[w];
                }
            }
            publish(F[0], F[RPT - 1], cur ^ 1);
            cur ^= 1;
            __syncthreads();
        };
        // level L even, before its step: plane ctz(L >> 1) += +-avail (nested sets: the alternating sum is the XOR)
        auto gray = [&](uint32_t L) {
            const uint32_t M = L >> 1;
            const int k = __ffs(M) - 1;
            const uint32_t s = ((M >> (k + 1)) & 1u) ? neg1 : one;
            if (k < WD_NPS) {
                uint32_t *pl = sm + k * PLANE;
#pragma unroll
                for (int r = 0; r < RPT; ++r) {
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
                for (int r = 0; r < RPT; ++r) {
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
        publish(F[0], F[RPT - 1], cur);
        __syncthreads();
        uint32_t L = 1;
        int cpar = 0;
        for (;; L += 4) {
#pragma unroll 1
            for (uint32_t h = 0; h < 4; h += 2) {
                step();
                gray(L + h + 1);
                step();
            }
            uint32_t any = 0;
#pragma unroll
            for (int r = 0; r < RPT; ++r)
#pragma unroll
                for (int w = 0; w < WD_W; ++w) any |= F[r][w];
            if constexpr (CL == 1) {
                if (!__syncthreads_or(any != 0)) break;
            } else {
                // cluster-wide vote: each CTA sends its own verdict to the peer (4 bytes on the peer's vote mbarrier) and waits for
                // the peer's; both CTAs vote at the same levels, so the phases stay paired
                const int mine = __syncthreads_or(any != 0);
                if (tid == 0) {
                    st_async4(peer_shared_u32(&conv_in[cpar], crank ^ 1u), static_cast<uint32_t>(mine), peer_shared_u32(mb_flag, crank ^ 1u) + 8 * cpar);
                    mbar_expect_tx(mb_flag32 + 8 * cpar, 4);
                }
                mbar_wait_cluster(mb_flag32 + 8 * cpar, cpar ? ph_flag1 : ph_flag0);
                if (cpar) ph_flag1 ^= 1u; else ph_flag0 ^= 1u;
                const bool more = (mine | static_cast<int>(*reinterpret_cast<volatile uint32_t *>(&conv_in[cpar]))) != 0;
                cpar ^= 1;
                if (!more) break;
            }
        }
        if constexpr (CL > 1) {
            if (seam_warp) seam_wait(cur);        // the peer's last row is never used: consumed here so that the phases stay paired
        }
        const uint32_t Mmax = (L + 2) >> 1;
        const int kmax = 32 - __clz(Mmax);

        // ---- 3. Gray -> binary in place (plane k becomes cost bit k + 1); the reached mask to the scratch ----
#pragma unroll 1
        for (int r = 0; r < RPT; ++r) {
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
                for (int k = kmax; k < 2; ++k) strow(sm + k * PLANE, r, tid, z);
            }
            ldrow(free_g, r, tid, fr);
#pragma unroll
            for (int u = 0; u < RPT; ++u)
                if (u == r) {
#pragma unroll
                    for (int w = 0; w < WD_W; ++w) v[w] = fr[w] & ~A[u][w];
                }
            strow(vis_g, r, tid, v);
        }
        if constexpr (CL > 1) __threadfence();     // the peer reads this CTA's reached mask of the seam row from the L2 scratch
        block_or_cluster_sync();

        // ---- 4. per row: flow direction -> flow bytes, cost planes -> int32 (word group by word group) ----
        uint8_t *flow = a.flow + plane * cells;
        int32_t *cost = a.cost ? a.cost + plane * cells : nullptr;
        const uint32_t *b1p = sm, *b2p = sm + PLANE;
#pragma unroll 1
        for (int r = 0; r < RPT; ++r) {
            const int R = RPT * gtid + r;
            // the rows above / below: the same thread's, a neighbouring thread's, or — at the seam of the cluster variant — a thread
            // of the peer CTA (its planes through DSMEM, its reached mask from its L2 scratch)
            const int gu = r == 0 ? gtid - 1 : gtid, ru = r == 0 ? RPT - 1 : r - 1;
            const int gd = r == RPT - 1 ? gtid + 1 : gtid, rd = r == RPT - 1 ? 0 : r + 1;
            const bool up_ok = gu >= 0, dn_ok = gd < CL * WD_T;
            const bool up_peer = CL > 1 && up_ok && (gu / WD_T) != static_cast<int>(crank);
            const bool dn_peer = CL > 1 && dn_ok && (gd / WD_T) != static_cast<int>(crank);
            const int tu = gu & (WD_T - 1), td = gd & (WD_T - 1);
            const uint32_t par0 = static_cast<uint32_t>(R + gi + gj) & 1u;
            rowops::RowInW<WD_W> in;
            ldrow(b1p, r, tid, in.b1c); ldrow(b2p, r, tid, in.b2c); ldrow(vis_g, r, tid, in.Vc); ldrow(free_g, r, tid, in.Fc);
#pragma unroll
            for (int w = 0; w < WD_W; ++w) { in.b1u[w] = in.b2u[w] = in.Vu[w] = 0u; in.b1d[w] = in.b2d[w] = in.Vd[w] = 0u; }
            if (up_ok) {
                ldrow(up_peer ? sm_peer : b1p, ru, tu, in.b1u); ldrow(up_peer ? sm_peer + PLANE : b2p, ru, tu, in.b2u);
                ldrow(up_peer ? vis_peer : vis_g, ru, tu, in.Vu);
            }
            if (dn_ok) {
                ldrow(dn_peer ? sm_peer : b1p, rd, td, in.b1d); ldrow(dn_peer ? sm_peer + PLANE : b2p, rd, td, in.b2d);
                ldrow(dn_peer ? vis_peer : vis_g, rd, td, in.Vd);
            }
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
    if (a.ticket && tid == 0) flow_launch_epilogue(a);
}

}  // namespace

bool flow_field_wide_supported(int G) { return G > 384 && G <= 512 && (G % 32) == 0; }
size_t flow_field_wide_scratch_words() { return static_cast<size_t>(WD_NPG + 2) * WD_PLANE; }
int flow_field_wide_max_grid() { return 148 * 2; }

// FFMP_FLOW_CLUSTER=1: grids of 384 < G <= 512 run on the cluster variant (two CTAs per grid).  Off by default: it is not
// faster; the parity suite runs both.
static bool wide_cluster_enabled() {
    const char *e = std::getenv("FFMP_FLOW_CLUSTER");
    return e && std::atoi(e) != 0;
}

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
            cudaError_t ce = cudaFuncSetAttribute(flow_field_wide_kernel<true, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, WD_SMEM);
            if (ce == cudaSuccess) ce = cudaFuncSetAttribute(flow_field_wide_kernel<false, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, WD_SMEM);
            if (ce == cudaSuccess) ce = cudaFuncSetAttribute(flow_field_wide_kernel<true, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, WD_SMEM_CL2);
            if (ce == cudaSuccess) ce = cudaFuncSetAttribute(flow_field_wide_kernel<false, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, WD_SMEM_CL2);
            if (ce != cudaSuccess) return ce;
            if (known) configured_dev[dev] = true;
        }
    }
    if (wide_cluster_enabled() && grid >= 2) {
        // one cluster of two CTAs per grid; `grid` is the CTA budget the caller's scratch was sized for: never more CTAs than that
        cudaLaunchConfig_t cfg = {};
        const int clusters = grid / 2;
        cfg.gridDim = dim3(static_cast<unsigned>(2 * clusters));
        cfg.blockDim = dim3(WD_T);
        cfg.dynamicSmemBytes = WD_SMEM_CL2;
        cfg.stream = st;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        if (a.generate) return cudaLaunchKernelEx(&cfg, flow_field_wide_kernel<true, 2>, a);
        return cudaLaunchKernelEx(&cfg, flow_field_wide_kernel<false, 2>, a);
    }
    if (a.generate) flow_field_wide_kernel<true, 1><<<grid, WD_T, WD_SMEM, st>>>(a);
    else flow_field_wide_kernel<false, 1><<<grid, WD_T, WD_SMEM, st>>>(a);
    return cudaGetLastError();
}

}  // namespace ffmp
