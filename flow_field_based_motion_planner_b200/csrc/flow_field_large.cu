// flow_field_large.cu — SPEC.md §4/§5 for large maps (128 < G <= 512, G % 32 == 0): one CTA per grid.
//
// The 128x128 path keeps a whole grid in the registers of one warp; a 512x512 grid (BASELINE config 4) is
// 32 KB per bit mask, so here the masks live in shared memory (avail, two frontier buffers, free:
// 4 x G^2/8 bytes = 128 KB at G=512) and 1024 threads sweep the words of the bit-parallel wavefront with
// one __syncthreads_or per level.  Words whose 3x3 word neighbourhood holds no frontier bit are skipped
// after one shared-memory probe.  The level of a cell is stored to the int32 cost plane when its bit
// first sets; the direction pass then reads the (L2-resident) cost plane.  First correct version: the
// per-level barrier makes it latency-bound; a cluster / DSMEM variant is the planned follow-up.
#include "ffmp_kernels.cuh"

namespace ffmp {

namespace {

constexpr int LARGE_THREADS = 1024;

struct LargeShared {
    unsigned long long plane;
    int gi, gj;
    uint32_t key;
    ScenarioParams sp;
};

template <bool GEN>
__global__ void __launch_bounds__(LARGE_THREADS) flow_field_cta_kernel(FlowArgs a) {
    extern __shared__ __align__(16) uint32_t sm[];
    __shared__ LargeShared sh;
    const int G = a.G;
    const int wpr = G >> 5;              // words per row
    const int nw = G * wpr;              // words per mask
    uint32_t *A = sm, *F0 = sm + nw, *F1 = sm + 2 * nw, *FR = sm + 3 * nw;
    const int tid = threadIdx.x;
    const int count = a.count_ptr ? static_cast<int>(*a.count_ptr) : a.count;
    const size_t cells = static_cast<size_t>(G) * G;

    for (int item = blockIdx.x; item < count; item += gridDim.x) {
        if (tid == 0) {
            const uint32_t env = a.env_idx ? a.env_idx[item] : static_cast<uint32_t>(item);
            if (GEN) {
                const uint32_t episode = a.episode ? a.episode[item] : a.episode_const;
                sh.plane = static_cast<unsigned long long>(episode % a.S) * a.N + env;
                sh.key = scenario_key(a.seed, a.env_id_base + env, episode);
                sh.sp = sample_scenario(sh.key, G, a.goal_mode);
                store_scenario_record(a.scen_out + sh.plane * SC_WORDS, sh.sp, sh.key);
                sh.gi = sh.sp.gi; sh.gj = sh.sp.gj;
            } else if (a.slot_mode) {
                sh.plane = static_cast<unsigned long long>((a.episode ? a.episode[item] : a.episode_const) % a.S) * a.N + env;
                sh.gi = static_cast<int>(a.scen[sh.plane * SC_WORDS + SC_GI]);
                sh.gj = static_cast<int>(a.scen[sh.plane * SC_WORDS + SC_GJ]);
            } else {
                sh.plane = static_cast<unsigned long long>(item);
                sh.gi = a.goal_cells[2 * item];
                sh.gj = a.goal_cells[2 * item + 1];
            }
        }
        __syncthreads();
        const size_t plane = static_cast<size_t>(sh.plane);
        const int gi = sh.gi, gj = sh.gj;
        int32_t *cost = a.cost + plane * cells;
        uint8_t *flow = a.flow + plane * cells;

        // ---- free mask (generated or packed from the occupancy bytes), cost plane pre-filled with INF ----
        for (int w = tid; w < nw; w += LARGE_THREADS) {
            const int R = w / wpr, wc = w - R * wpr;
            uint32_t fr;
            if (GEN) {
                fr = scenario_free_word(sh.key, R, 32 * wc, G, a.block_shift, a.p_thresh, sh.sp);
            } else {
                const uint4 *src = reinterpret_cast<const uint4 *>(a.occ + plane * cells + static_cast<size_t>(R) * G + 32 * wc);
                fr = 0;
#pragma unroll
                for (int q = 0; q < 2; ++q) {
                    const uint4 v = __ldg(src + q);
                    const uint32_t x4[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const uint32_t x = x4[u];
                        const uint32_t nz = (x | ((x & 0x7F7F7F7Fu) + 0x7F7F7F7Fu)) & 0x80808080u;
                        const uint32_t f = (~nz >> 7) & 0x01010101u;
                        fr |= ((f * 0x01020408u) >> 24 & 0xFu) << (16 * q + 4 * u);
                    }
                }
            }
            FR[w] = fr; A[w] = fr; F0[w] = 0; F1[w] = 0;
        }
        {
            int4 *c4 = reinterpret_cast<int4 *>(cost);
            const int4 inf = make_int4(COST_INF, COST_INF, COST_INF, COST_INF);
            for (size_t c = tid; c < cells / 4; c += LARGE_THREADS) c4[c] = inf;
        }
        __syncthreads();
        if (tid == 0 && gi >= 0 && gj >= 0 && gi < G && gj < G) {
            const int w = gi * wpr + (gj >> 5);
            const uint32_t m = (1u << (gj & 31)) & A[w];
            if (m) { F0[w] = m; A[w] &= ~m; cost[static_cast<size_t>(gi) * G + gj] = 0; }
        }
        __syncthreads();

        // ---- level-synchronous bit-parallel wavefront ----------------------------------------------
        uint32_t *F = F0, *Fn = F1;
        for (int L = 1;; ++L) {
            int any = 0;
            for (int w = tid; w < nw; w += LARGE_THREADS) {
                const int R = w / wpr, wc = w - R * wpr;
                const uint32_t c = F[w];
                const uint32_t lw = wc > 0 ? F[w - 1] : 0u, rw = wc < wpr - 1 ? F[w + 1] : 0u;
                const uint32_t up = R > 0 ? F[w - wpr] : 0u, dn = R < G - 1 ? F[w + wpr] : 0u;
                uint32_t n = 0;
                if (c | lw | rw | up | dn) n = ((c << 1) | (lw >> 31) | (c >> 1) | (rw << 31) | up | dn) & A[w];
                Fn[w] = n;
                if (n) {
                    A[w] &= ~n;
                    any = 1;
                    int32_t *dst = cost + static_cast<size_t>(R) * G + 32 * wc;
                    uint32_t m = n;
                    while (m) {
                        const int b = __ffs(m) - 1;
                        dst[b] = L;
                        m &= m - 1;
                    }
                }
            }
            if (!__syncthreads_or(any)) break;
            uint32_t *t = F; F = Fn; Fn = t;
        }
        __syncthreads();   // cost stores of every thread are visible to the CTA

        // ---- flow direction (explicit 8-neighbour scan on the cost plane) and flow image -----------
        for (size_t c = tid; c < cells; c += LARGE_THREADS) {
            const int i = static_cast<int>(c / G), j = static_cast<int>(c - static_cast<size_t>(i) * G);
            const bool occupied = !((FR[i * wpr + (j >> 5)] >> (j & 31)) & 1u);
            uint32_t d = 8;
            const int own = cost[c];
            if (own != COST_INF) {
                auto at = [&](int ii, int jj) -> int {
                    return (ii < 0 || jj < 0 || ii >= G || jj >= G) ? COST_INF : cost[static_cast<size_t>(ii) * G + jj];
                };
                const int cE = at(i + 1, j), cN = at(i, j + 1), cW = at(i - 1, j), cS = at(i, j - 1);
                int best = own;
                // scan order E, NE, N, NW, W, SW, S, SE with strict '<'; a diagonal needs both side cells free
                // (a free side cell next to a reached cell is reached, so "free" == "cost != INF" here)
                if (cE < best) { best = cE; d = 0; }
                if (cE != COST_INF && cN != COST_INF) { const int v = at(i + 1, j + 1); if (v < best) { best = v; d = 1; } }
                if (cN < best) { best = cN; d = 2; }
                if (cW != COST_INF && cN != COST_INF) { const int v = at(i - 1, j + 1); if (v < best) { best = v; d = 3; } }
                if (cW < best) { best = cW; d = 4; }
                if (cW != COST_INF && cS != COST_INF) { const int v = at(i - 1, j - 1); if (v < best) { best = v; d = 5; } }
                if (cS < best) { best = cS; d = 6; }
                if (cE != COST_INF && cS != COST_INF) { const int v = at(i + 1, j - 1); if (v < best) { best = v; d = 7; } }
            }
            flow[c] = occupied ? 255 : static_cast<uint8_t>(d * 28);
        }
        __syncthreads();
    }

    if (a.ticket && tid == 0) {
        __threadfence();
        const uint32_t t = atomicAdd(a.ticket, 1u);
        if (t == gridDim.x - 1) {
            *a.ticket = 0;
            if (a.count_reset) *a.count_reset = 0;
            __threadfence();
        }
    }
}

size_t large_smem_bytes(int G) { return static_cast<size_t>(4) * G * (G >> 5) * sizeof(uint32_t); }

}  // namespace

bool flow_field_large_supported(int G) { return G > 128 && G <= 512 && (G % 32) == 0; }

int flow_field_large_max_grid(int G) {
    int per_sm = static_cast<int>((227 * 1024) / (large_smem_bytes(G) + 1024));
    if (per_sm < 1) per_sm = 1;
    if (per_sm > 2) per_sm = 2;   // 1024 threads per CTA
    return 148 * per_sm;
}

cudaError_t launch_flow_field_large(const FlowArgs &a, int grid, cudaStream_t st) {
    if (grid <= 0) return cudaSuccess;
    if (!a.cost) return cudaErrorInvalidValue;   // the direction pass reads the cost plane
    const size_t smem = large_smem_bytes(a.G);
    static bool configured = false;
    if (!configured) {
        cudaError_t ce = cudaFuncSetAttribute(flow_field_cta_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        if (ce == cudaSuccess)
            ce = cudaFuncSetAttribute(flow_field_cta_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        if (ce != cudaSuccess) return ce;
        configured = true;
    }
    if (a.generate) flow_field_cta_kernel<true><<<grid, LARGE_THREADS, smem, st>>>(a);
    else flow_field_cta_kernel<false><<<grid, LARGE_THREADS, smem, st>>>(a);
    return cudaGetLastError();
}

}  // namespace ffmp
