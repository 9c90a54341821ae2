// flow_field_large.cu — SPEC.md §4/§5 for large maps (128 < G <= 512, G % 32 == 0): one CTA per grid, one THREAD per row.
//
// The 128x128 kernel keeps a whole grid in the registers of one warp.  A 512x512 grid is 32 KB per bit mask, so here a
// CTA of P threads (P = 256 or 512 padded rows) owns the grid: thread t holds row t of the avail / frontier masks as WPR
// 32-bit words in REGISTERS (16 words at G = 512), exactly the per-lane layout of the small kernel with one row per lane,
// and the same ALU-lean wavefront step (funnel shifts + two LOP3 per word, avail update on the FMA pipe).  Rows above /
// below come from a double-buffered shared-memory frontier buffer (chunk-major, conflict-free 16-byte accesses), and one
// __syncthreads_or per level both orders that exchange and tests convergence.  BFS levels are recorded as Gray-code
// bit-planes: plane 0 in registers, planes 1..NPSL in shared memory, the rest (touched every 2^(k+1) levels) in a per-CTA
// global scratch that stays in L2.  Afterwards each thread un-Grays its row, evaluates the 8-neighbour argmin bit-parallel
// (neighbour rows from the shared-memory planes), and writes its row of the flow image and of the int32 integration field:
// one word of the mask is exactly one 128-byte line of cost, so the per-thread 16-byte stores fill whole lines.
// Algorithmic HBM bytes: 6 B/cell (1 occ read + 4 cost write + 1 flow write).
#include <mutex>
#include "flow_bits.cuh"

namespace ffmp {

namespace {

constexpr int NPSL = 3;               // shared-memory Gray planes 1..NPSL (plane 0 lives in registers)
constexpr int NPMAX = 18;             // cost bits: depth < 2^18 = 512 * 512
constexpr int NPGL = NPMAX - 1 - NPSL;   // planes NPSL+1..17 in the global scratch
constexpr int NSM = NPSL + 2;         // shared-memory planes per CTA: 1..NPSL, visited, free

struct LargeInfo {
    int item;
    unsigned long long plane;
    int gi, gj;
    uint32_t key;
    ScenarioParams sp;
};

// 16 occupancy bytes -> 16 free bits (bit = 1: byte == 0)
__device__ __forceinline__ uint32_t pack16(const uint4 q) {
    const uint32_t x4[4] = {q.x, q.y, q.z, q.w};
    uint32_t bits = 0;
#pragma unroll
    for (int u = 0; u < 4; ++u) {
        const uint32_t x = x4[u];
        const uint32_t nz = (x | ((x & 0x7F7F7F7Fu) + 0x7F7F7F7Fu)) & 0x80808080u;   // 0x80 per non-zero byte
        const uint32_t fr = (~nz >> 7) & 0x01010101u;
        bits |= ((fr * 0x01020408u) >> 24 & 0xFu) << (4 * u);
    }
    return bits;
}

template <int WPR, bool GEN>
__global__ void __launch_bounds__(32 * WPR, 1) flow_field_rows_kernel(FlowArgs a) {
    constexpr int P = 32 * WPR;                 // padded grid side = threads per CTA
    constexpr int NQ = WPR / 4;                 // 16-byte chunks per row
    constexpr int PLANE_WORDS = P * WPR;        // one bit-plane of the padded grid
    constexpr int FBUF_WORDS = (P + 2) * WPR;   // a frontier buffer: rows -1..P (the two ghost rows stay zero)
    // Shared-memory planes are CHUNK-MAJOR, word (r, w) at ((w / 4) * rows + r) * 4 + w % 4, so that the 16-byte accesses of
    // consecutive lanes (consecutive rows) are consecutive in memory (a row-major layout would be a 4-way bank conflict).
    //   [0, NPSL)            Gray planes 1..NPSL
    //   then two frontier buffers (all rows, double buffered) during the BFS; afterwards the same storage holds the
    //   reached (pv) and free (pf) masks for the direction pass.
    extern __shared__ __align__(16) uint32_t pls[];
    __shared__ LargeInfo info;
    uint32_t *const fb0 = pls + NPSL * PLANE_WORDS;
    uint32_t *const fb1 = fb0 + FBUF_WORDS;
    uint32_t *const pv = fb0, *const pf = fb1;
    // activity masks of the published frontier rows (double buffered like the rows, P + 2 entries each incl. the ghost rows):
    //   bits 0..3  chunk q (words 4q..4q+3) of the row is non-zero
    //   bits 4..7  bit 31 of the chunk's last word is set  (the row's chunk q+1 sees it as its west neighbour)
    //   bits 8..11 bit 0 of the chunk's first word is set  (the row's chunk q-1 sees it as its east neighbour)
    uint16_t *const am0 = reinterpret_cast<uint16_t *>(fb1 + FBUF_WORDS);
    uint16_t *const am1 = am0 + (P + 2);
    constexpr uint32_t NZ_MASK = (1u << NQ) - 1u;

    const int tid = threadIdx.x;
    const int row = tid;
    const int G = a.G;
    const int count = a.count_ptr ? static_cast<int>(*a.count_ptr) : a.count;
    // per-CTA global scratch (L2): Gray planes NPSL+1..NPMAX-1 and the free mask, row-major [row][WPR]
    uint32_t *hi = a.hi_scratch + static_cast<size_t>(blockIdx.x) * ((NPGL + 1) * PLANE_WORDS);
    uint32_t *const free_g = hi + NPGL * PLANE_WORDS + row * WPR;
    const size_t cells = static_cast<size_t>(G) * G;
    const uint32_t neg1 = a.neg1;
    auto gl_plane = [&](int k) -> uint32_t * { return &hi[((k - 1 - NPSL) * P + row) * WPR]; };           // k > NPSL, own row
    auto sm_ld = [&](const uint32_t *base, int rows, int r, uint32_t (&v)[WPR]) {
#pragma unroll
        for (int q = 0; q < NQ; ++q) {
            const uint4 t = *reinterpret_cast<const uint4 *>(base + (q * rows + r) * 4);
            v[4 * q] = t.x; v[4 * q + 1] = t.y; v[4 * q + 2] = t.z; v[4 * q + 3] = t.w;
        }
    };
    auto sm_st = [&](uint32_t *base, int rows, int r, const uint32_t (&v)[WPR]) {
#pragma unroll
        for (int q = 0; q < NQ; ++q)
            *reinterpret_cast<uint4 *>(base + (q * rows + r) * 4) = make_uint4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
    };

    // grids are handed out through a device counter (a.work) when there is one: depths differ by 2x between maps, and 512
    // grids over 148 CTAs would otherwise be four static rounds for every CTA that drew item 444..511
    for (int item = blockIdx.x;; item += gridDim.x) {
        if (a.work) {
            if (tid == 0) info.item = static_cast<int>(atomicAdd(a.work, 1u));
            __syncthreads();
            item = info.item;
        }
        if (item >= count) break;
        // ---- 0. item parameters ------------------------------------------------------------------------
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

        // ---- 1. free-cell mask of this thread's row; frontier seed; zeroed planes -------------------------
        uint32_t A[WPR], F[WPR], G0[WPR];
        if (GEN) {
#pragma unroll 1
            for (int w = 0; w < WPR; ++w) {
                const uint32_t v = scenario_free_word(info.key, row, 32 * w, G, a.block_shift, a.p_thresh, info.sp);
#pragma unroll
                for (int u = 0; u < WPR; ++u)
                    if (u == w) A[u] = v;     // static register index
            }
        } else {
            // the thread's row is G contiguous bytes (16-byte aligned: G % 32 == 0): two 16-byte loads per mask word
            const uint4 *src = reinterpret_cast<const uint4 *>(a.occ + plane * cells + static_cast<size_t>(row) * G);
#pragma unroll
            for (int w = 0; w < WPR; ++w) {
                uint32_t v = 0;
                if (row < G && 32 * w < G) v = pack16(__ldg(src + 2 * w)) | (pack16(__ldg(src + 2 * w + 1)) << 16);
                A[w] = v;
            }
        }
        Row<WPR>::st(free_g, A);              // kept for the post-BFS phases; the loop below only needs `avail`
        {
            uint32_t z[WPR];
            const int gw = gj >> 5;
            const uint32_t bit = (row == gi && gj >= 0 && gj < G) ? (1u << (gj & 31)) : 0u;
#pragma unroll
            for (int w = 0; w < WPR; ++w) {
                z[w] = 0; G0[w] = 0;
                const uint32_t m = (w == gw) ? (bit & A[w]) : 0u;
                F[w] = m;
                A[w] ^= m;
            }
#pragma unroll
            for (int k = 0; k < NPSL; ++k) sm_st(pls + k * PLANE_WORDS, P, row, z);
            if (tid < 2) {            // ghost rows -1 and P of both frontier buffers
                sm_st(fb0, P + 2, tid * (P + 1), z);
                sm_st(fb1, P + 2, tid * (P + 1), z);
                am0[tid * (P + 1)] = 0; am1[tid * (P + 1)] = 0;
            }
        }
        sm_st(fb0, P + 2, row + 1, F);
        uint32_t m_own = 0;           // activity mask of this thread's frontier row (see above)
#pragma unroll
        for (int q = 0; q < NQ; ++q) {
            const uint32_t nz = F[4 * q] | F[4 * q + 1] | F[4 * q + 2] | F[4 * q + 3];
            m_own |= (nz != 0 ? 1u : 0u) << q | (F[4 * q + 3] >> 31) << (4 + q) | (F[4 * q] & 1u) << (8 + q);
        }
        am0[row + 1] = static_cast<uint16_t>(m_own);
        __syncthreads();

        // ---- 2. bit-parallel wavefront -----------------------------------------------------------------
        // Every level: the rows above / below come from the frontier buffer published by the previous level (two vector
        // loads per 4 words, conflict-free), the words are walked in place (funnel shifts use the old F[w]), the new
        // frontier row is published into the other buffer, and one __syncthreads_or both orders the exchange and tests
        // convergence.
        uint32_t *fcur = fb0, *fnext = fb1;
        uint16_t *acur = am0, *anext = am1;
        uint32_t L = 1;
        for (;; ++L) {
            // Gray bit-plane update: cells with cost >= L flip Gray bit ctz(L)
            const int k = __ffs(L) - 1;
            if (k == 0) {
#pragma unroll
                for (int w = 0; w < WPR; ++w) G0[w] ^= A[w];
            } else if (k <= NPSL) {
                uint32_t v[WPR];
                uint32_t *p = pls + (k - 1) * PLANE_WORDS;
                sm_ld(p, P, row, v);
#pragma unroll
                for (int w = 0; w < WPR; ++w) v[w] ^= A[w];
                sm_st(p, P, row, v);
            } else {
                const bool first = L == (1u << k);
                uint32_t *p = gl_plane(k);
                uint32_t v[WPR];
                if (first) {
#pragma unroll
                    for (int w = 0; w < WPR; ++w) v[w] = A[w];
                } else {
                    Row<WPR>::ld(p, v);
#pragma unroll
                    for (int w = 0; w < WPR; ++w) v[w] ^= A[w];
                }
                Row<WPR>::st(p, v);
            }
            // Which chunks can receive new cells at this level?  Those that hold frontier cells in this row or in the rows above /
            // below, plus the horizontal neighbours of an edge bit.  OR-ed over the warp (one REDUX) the answer is warp-uniform:
            // a skipped chunk costs nothing, and its frontier words are zero before and after (they were part of the test).
            uint32_t act = (m_own | acur[row] | acur[row + 2]) & NZ_MASK;
            act |= (((m_own >> 4) & NZ_MASK) << 1) | (((m_own >> 8) & NZ_MASK) >> 1);
            const uint32_t wact = __reduce_or_sync(FULL, act & NZ_MASK);
            uint32_t any = 0, f_prev = 0, m_new = 0;
#pragma unroll
            for (int q = 0; q < NQ; ++q) {
                if (wact & (1u << q)) {
                    const uint4 u4 = *reinterpret_cast<const uint4 *>(fcur + (q * (P + 2) + row) * 4);        // grid row - 1
                    const uint4 d4 = *reinterpret_cast<const uint4 *>(fcur + (q * (P + 2) + row + 2) * 4);    // grid row + 1
                    const uint32_t up[4] = {u4.x, u4.y, u4.z, u4.w}, dn[4] = {d4.x, d4.y, d4.z, d4.w};
                    uint32_t nz = 0;
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int w = 4 * q + j;
                        const uint32_t f_cur = F[w], f_next = w + 1 < WPR ? F[w + 1] : 0u;
                        const uint32_t lo = w > 0 ? __funnelshift_l(f_prev, f_cur, 1) : f_cur << 1;
                        const uint32_t hv = w + 1 < WPR ? __funnelshift_r(f_cur, f_next, 1) : f_cur >> 1;
                        const uint32_t n = (lo | hv | up[j] | dn[j]) & A[w];
                        nz |= n;
                        A[w] = sub_on_fma(A[w], n, neg1);
                        F[w] = n;
                        f_prev = f_cur;
                    }
                    any |= nz;
                    m_new |= (nz != 0 ? 1u : 0u) << q | (F[4 * q + 3] >> 31) << (4 + q) | (F[4 * q] & 1u) << (8 + q);
                } else {
                    f_prev = 0;       // == the (zero) last frontier word of the skipped chunk
                }
            }
            sm_st(fnext, P + 2, row + 1, F);
            anext[row + 1] = static_cast<uint16_t>(m_new);
            m_own = m_new;
            if (!__syncthreads_or(any != 0)) break;
            uint32_t *t = fcur; fcur = fnext; fnext = t;
            uint16_t *ta = acur; acur = anext; anext = ta;
        }
        const uint32_t Lmax = L - 1;                       // deepest level that reached a cell
        const int kmax = 32 - __clz(Lmax);                  // number of significant cost bits (<= NPMAX)

        // ---- 3. Gray -> binary of this thread's row (planes stay where they live); reached / free rows to shared memory ----
        uint32_t B0[WPR];
        {
            uint32_t acc[WPR];
#pragma unroll
            for (int w = 0; w < WPR; ++w) acc[w] = 0;
#pragma unroll 1
            for (int k = kmax - 1; k > NPSL; --k) {
                uint32_t *p = gl_plane(k);
                uint32_t v[WPR];
                Row<WPR>::ld(p, v);
#pragma unroll
                for (int w = 0; w < WPR; ++w) { acc[w] ^= v[w]; v[w] = acc[w]; }
                Row<WPR>::st(p, v);
            }
#pragma unroll 1
            for (int k = min(kmax - 1, NPSL); k >= 1; --k) {
                uint32_t *p = pls + (k - 1) * PLANE_WORDS;
                uint32_t v[WPR];
                sm_ld(p, P, row, v);
#pragma unroll
                for (int w = 0; w < WPR; ++w) { acc[w] ^= v[w]; v[w] = acc[w]; }
                sm_st(p, P, row, v);
            }
            uint32_t V[WPR], FRr[WPR];
            Row<WPR>::ld(free_g, FRr);
#pragma unroll
            for (int w = 0; w < WPR; ++w) { B0[w] = kmax > 0 ? (acc[w] ^ G0[w]) : 0u; V[w] = FRr[w] & ~A[w]; }
            sm_st(pv, P, row, V);     // the frontier buffers are dead: every thread left the loop through the barrier
            sm_st(pf, P, row, FRr);
        }
        __syncthreads();

        // ---- 4. flow direction, bit-parallel (SPEC.md §4 F2), and the flow image ------------------------
        // Everything but this row's cost bit 0 is read from shared memory, word by word (rows R-1 "u" = west, R+1 "d" = east).
        {
            const uint32_t *p1 = pls, *p2 = pls + PLANE_WORDS;       // binary cost bits 1 and 2
            const bool has1 = kmax > 1, has2 = kmax > 2;
            auto ldw = [&](const uint32_t *pl_, bool on, int r, int w) -> uint32_t {
                return (on && r >= 0 && r < P && w >= 0 && w < WPR) ? pl_[((w >> 2) * P + r) * 4 + (w & 3)] : 0u;
            };
            // value of the cell at column+1 / column-1 aligned to word w of row r
            auto hiw = [&](const uint32_t *pl_, bool on, int r, int w) { return (ldw(pl_, on, r, w) >> 1) | (ldw(pl_, on, r, w + 1) << 31); };
            auto low = [&](const uint32_t *pl_, bool on, int r, int w) { return (ldw(pl_, on, r, w) << 1) | (ldw(pl_, on, r, w - 1) >> 31); };
            uint8_t *flow = a.flow + plane * cells + static_cast<size_t>(row) * G;
#pragma unroll 1
            for (int w = 0; w < WPR; ++w) {
                if (row >= G || 32 * w >= G) break;
                uint32_t b0 = 0;
#pragma unroll
                for (int u = 0; u < WPR; ++u)
                    if (u == w) b0 = B0[u];
                const int ru = row - 1, rd = row + 1;
                const uint32_t b1c = ldw(p1, has1, row, w), b2c = ldw(p2, has2, row, w);
                const uint32_t own = ldw(pv, true, row, w);
                const uint32_t fc = ldw(pf, true, row, w);
                const uint32_t t = b1c ^ ~b0;    // bit 1 of (cost-1)
                const uint32_t u2 = b2c ^ ~b1c;  // bit 2 of (cost-2)
                // orthogonal neighbours one level lower (codes 0 E, 2 N, 4 W, 6 S)
                const uint32_t lE = own & ldw(pv, true, rd, w) & ~(ldw(p1, has1, rd, w) ^ t);
                const uint32_t lW = own & ldw(pv, true, ru, w) & ~(ldw(p1, has1, ru, w) ^ t);
                const uint32_t lN = own & hiw(pv, true, row, w) & ~(hiw(p1, has1, row, w) ^ t);
                const uint32_t lS = own & low(pv, true, row, w) & ~(low(p1, has1, row, w) ^ t);
                // admissible diagonals two levels lower (codes 1 NE, 3 NW, 5 SW, 7 SE)
                const uint32_t fE = ldw(pf, true, rd, w), fW = ldw(pf, true, ru, w), fN = hiw(pf, true, row, w), fS = low(pf, true, row, w);
                const uint32_t lNE = own & hiw(pf, true, rd, w) & fE & fN & (hiw(p1, has1, rd, w) ^ b1c) & ~(hiw(p2, has2, rd, w) ^ u2);
                const uint32_t lNW = own & hiw(pf, true, ru, w) & fW & fN & (hiw(p1, has1, ru, w) ^ b1c) & ~(hiw(p2, has2, ru, w) ^ u2);
                const uint32_t lSW = own & low(pf, true, ru, w) & fW & fS & (low(p1, has1, ru, w) ^ b1c) & ~(low(p2, has2, ru, w) ^ u2);
                const uint32_t lSE = own & low(pf, true, rd, w) & fE & fS & (low(p1, has1, rd, w) ^ b1c) & ~(low(p2, has2, rd, w) ^ u2);
                const uint32_t anyD = lNE | lNW | lSW | lSE;
                const uint32_t m0 = (anyD & lNE) | (~anyD & lE);
                const uint32_t m1 = (anyD & lNW) | (~anyD & lN);
                const uint32_t m2 = (anyD & lSW) | (~anyD & lW);
                const uint32_t m3 = (anyD & lSE) | (~anyD & lS);
                const uint32_t d1 = ~m0 & (m1 | (~m2 & m3));
                const uint32_t d2 = ~m0 & ~m1 & (m2 | m3);
                const uint32_t d3 = ~(m0 | m1 | m2 | m3);
                uint32_t out[8];
                flow_bytes32(anyD, d1, d2, d3, ~fc, out);
                uint4 *dst = reinterpret_cast<uint4 *>(flow + 32 * w);
                dst[0] = make_uint4(out[0], out[1], out[2], out[3]);
                dst[1] = make_uint4(out[4], out[5], out[6], out[7]);
            }
        }

        // ---- 5. integration field: bit-planes -> int32 (INF for cells not reached) -----------------------
        // One mask word is one 128-byte line of cost: the thread writes it with eight 16-byte stores.
        if (row < G && a.cost) {
            int32_t *cost = a.cost + plane * cells + static_cast<size_t>(row) * G;
#pragma unroll 1
            for (int w = 0; w < WPR; ++w) {
                if (32 * w >= G) break;
                uint32_t bw[NPMAX];
#pragma unroll
                for (int k = 0; k < NPMAX; ++k) bw[k] = 0;
                const uint32_t vis = pv[((w >> 2) * P + row) * 4 + (w & 3)];
#pragma unroll
                for (int u = 0; u < WPR; ++u)
                    if (u == w) bw[0] = B0[u];
#pragma unroll
                for (int k = 1; k < NPMAX; ++k)
                    if (k < kmax) bw[k] = k <= NPSL ? pls[(k - 1) * PLANE_WORDS + ((w >> 2) * P + row) * 4 + (w & 3)] : gl_plane(k)[w];
                uint32_t lo8[8], hi8[8], top8[8];   // cost bits 0-7 / 8-15 / 16-17 of the 32 cells, 4 cells per word
                {
                    uint32_t tl[4], th[4];
                    bytes4x4(bw[0], bw[1], bw[2], bw[3], tl);
                    bytes4x4(bw[4], bw[5], bw[6], bw[7], th);
#pragma unroll
                    for (int b = 0; b < 4; ++b) {
                        uint32_t x = tl[b], y = th[b];
                        transpose8(x, y);
                        lo8[2 * b] = x; lo8[2 * b + 1] = y;
                    }
                }
#pragma unroll
                for (int b = 0; b < 8; ++b) { hi8[b] = 0; top8[b] = 0; }
                if (kmax > 8) {
                    uint32_t tl[4], th[4];
                    bytes4x4(bw[8], bw[9], bw[10], bw[11], tl);
                    bytes4x4(bw[12], bw[13], bw[14], bw[15], th);
#pragma unroll
                    for (int b = 0; b < 4; ++b) {
                        uint32_t x = tl[b], y = th[b];
                        transpose8(x, y);
                        hi8[2 * b] = x; hi8[2 * b + 1] = y;
                    }
                }
                if (kmax > 16) {
                    uint32_t tl[4];
                    bytes4x4(bw[16], bw[17], 0u, 0u, tl);
#pragma unroll
                    for (int b = 0; b < 4; ++b) {
                        uint32_t x = tl[b], y = 0u;
                        transpose8(x, y);
                        top8[2 * b] = x; top8[2 * b + 1] = y;
                    }
                }
                int4 *dst = reinterpret_cast<int4 *>(cost + 32 * w);
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const uint32_t l4 = lo8[q], h4 = hi8[q], t4 = top8[q], vb = vis >> (4 * q);
                    int4 c;
                    c.x = (vb & 1u) ? static_cast<int>((l4 & 0xFFu) | ((h4 & 0xFFu) << 8) | ((t4 & 0xFFu) << 16)) : COST_INF;
                    c.y = (vb & 2u) ? static_cast<int>(((l4 >> 8) & 0xFFu) | (((h4 >> 8) & 0xFFu) << 8) | (((t4 >> 8) & 0xFFu) << 16)) : COST_INF;
                    c.z = (vb & 4u) ? static_cast<int>(((l4 >> 16) & 0xFFu) | (((h4 >> 16) & 0xFFu) << 8) | (((t4 >> 16) & 0xFFu) << 16)) : COST_INF;
                    c.w = (vb & 8u) ? static_cast<int>((l4 >> 24) | ((h4 >> 24) << 8) | ((t4 >> 24) << 16)) : COST_INF;
                    dst[q] = c;
                }
            }
        }
        __syncthreads();   // the planes are reused by the next grid
    }

    // the last CTA to finish re-arms the regeneration list for its next use
    if (a.ticket && tid == 0) flow_launch_epilogue(a);
}

int large_wpr(int G) { return G <= 128 ? 4 : G <= 256 ? 8 : 16; }
size_t large_smem_bytes(int G) {
    const int wpr = large_wpr(G);
    return (static_cast<size_t>(NPSL) * 32 * wpr * wpr + 2 * static_cast<size_t>(32 * wpr + 2) * wpr) * sizeof(uint32_t) +
           2 * static_cast<size_t>(32 * wpr + 2) * sizeof(uint16_t) + 8;      // + the activity masks of the frontier rows
}

}  // namespace

bool flow_field_large_supported(int G) { return G > 128 && G <= 512 && (G % 32) == 0; }
bool flow_field_rows_usable(int G) { return G >= 32 && G <= 512 && (G % 32) == 0; }

size_t flow_field_large_scratch_words(int G) {
    const int wpr = large_wpr(G);
    return static_cast<size_t>(NPGL + 1) * 32 * wpr * wpr;    // Gray planes NPSL+1.. and the free mask
}

int flow_field_large_max_grid(int G) {
    if (large_wpr(G) == 16) return 148;                       // 512 threads, 160 KB of planes: one CTA per SM
    int per_sm = static_cast<int>((227 * 1024) / (large_smem_bytes(G) + 2 * 1024));
    const int reg_limit = large_wpr(G) == 8 ? 2 : 5;          // ~96 registers x 256 / 128 threads
    if (per_sm > reg_limit) per_sm = reg_limit;
    return 148 * per_sm;
}

cudaError_t launch_flow_field_large(const FlowArgs &a_in, int grid, cudaStream_t st) {
    if (grid <= 0) return cudaSuccess;
    FlowArgs a = a_in;
    a.neg1 = 0xFFFFFFFFu;
    a.one = 1u;
    const size_t smem = large_smem_bytes(a.G);
    // the opt-in belongs to the device (context) of the call: tracked per ordinal (one process may drive several GPUs)
    static std::mutex mu;
    static bool configured_dev[64] = {false};
    int dev = 0;
    if (cudaError_t ce = cudaGetDevice(&dev); ce != cudaSuccess) return ce;
    std::lock_guard<std::mutex> lock(mu);
    const bool known = dev >= 0 && dev < 64;
    if (!known || !configured_dev[dev]) {
        cudaError_t ce = cudaFuncSetAttribute(flow_field_rows_kernel<16, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        if (ce == cudaSuccess) ce = cudaFuncSetAttribute(flow_field_rows_kernel<16, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        if (ce == cudaSuccess) ce = cudaFuncSetAttribute(flow_field_rows_kernel<8, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
        if (ce == cudaSuccess) ce = cudaFuncSetAttribute(flow_field_rows_kernel<8, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
        if (ce == cudaSuccess) ce = cudaFuncSetAttribute(flow_field_rows_kernel<4, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 48 * 1024);
        if (ce == cudaSuccess) ce = cudaFuncSetAttribute(flow_field_rows_kernel<4, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 48 * 1024);
        if (ce != cudaSuccess) return ce;
        if (known) configured_dev[dev] = true;
    }
    if (large_wpr(a.G) == 16) {
        if (a.generate) flow_field_rows_kernel<16, true><<<grid, 512, smem, st>>>(a);
        else flow_field_rows_kernel<16, false><<<grid, 512, smem, st>>>(a);
    } else if (large_wpr(a.G) == 4) {
        if (a.generate) flow_field_rows_kernel<4, true><<<grid, 128, smem, st>>>(a);
        else flow_field_rows_kernel<4, false><<<grid, 128, smem, st>>>(a);
    } else {
        if (a.generate) flow_field_rows_kernel<8, true><<<grid, 256, smem, st>>>(a);
        else flow_field_rows_kernel<8, false><<<grid, 256, smem, st>>>(a);
    }
    return cudaGetLastError();
}

}  // namespace ffmp
