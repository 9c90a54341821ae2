// flow_order.cu — hand-out order of a flow-field batch: deepest grids first.
// The flow-field kernels process one grid per warp (or CTA) and hand grids out through a device counter; a batch of a few
// grids per warp ends with a tail in which the SMs drain one after the other.  The number of BFS levels of a grid is close to
// the Manhattan distance from the goal to the farthest corner, so a counting sort of the batch by that estimate (longest
// processing time first) shortens the tail.  The order only changes WHEN a grid is computed, never its result.
#include "ffmp_kernels.cuh"

namespace ffmp {

namespace {

constexpr int ORDER_THREADS = 1024;
constexpr int ORDER_KEYS = 2048;       // estimates 0 .. 2 * 1024 - 2

__global__ void __launch_bounds__(ORDER_THREADS) flow_order_kernel(const int32_t *goal_cells, int n, int G, uint32_t *order) {
    __shared__ uint32_t hist[ORDER_KEYS];
    __shared__ uint32_t warp_tot[ORDER_THREADS / 32];
    const int tid = threadIdx.x;
    for (int k = tid; k < ORDER_KEYS; k += ORDER_THREADS) hist[k] = 0;
    __syncthreads();
    auto key_of = [&](int i) {
        const int gi = goal_cells[2 * i], gj = goal_cells[2 * i + 1];
        int est = 0;
        if (gi >= 0 && gj >= 0 && gi < G && gj < G) est = max(gi, G - 1 - gi) + max(gj, G - 1 - gj);
        return 2 * G - 2 - est;        // ascending key = descending depth estimate
    };
    for (int i = tid; i < n; i += ORDER_THREADS) atomicAdd(&hist[key_of(i)], 1u);
    __syncthreads();
    // exclusive prefix sum over the 2048 keys: two keys per thread, warp scans, one scan of the warp totals
    const uint32_t a0 = hist[2 * tid], a1 = hist[2 * tid + 1];
    uint32_t v = a0 + a1;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t t = __shfl_up_sync(FULL, v, d);
        if ((tid & 31) >= d) v += t;
    }
    if ((tid & 31) == 31) warp_tot[tid >> 5] = v;
    __syncthreads();
    if (tid < 32) {
        uint32_t t = warp_tot[tid];
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t u = __shfl_up_sync(FULL, t, d);
            if (tid >= d) t += u;
        }
        warp_tot[tid] = t;
    }
    __syncthreads();
    const uint32_t base = v - (a0 + a1) + ((tid >> 5) ? warp_tot[(tid >> 5) - 1] : 0u);
    hist[2 * tid] = base;
    hist[2 * tid + 1] = base + a0;
    __syncthreads();
    for (int i = tid; i < n; i += ORDER_THREADS) order[atomicAdd(&hist[key_of(i)], 1u)] = static_cast<uint32_t>(i);
}

}  // namespace

cudaError_t launch_flow_order(const int32_t *goal_cells, int n, int G, uint32_t *order, cudaStream_t st) {
    if (n <= 0) return cudaSuccess;
    flow_order_kernel<<<1, ORDER_THREADS, 0, st>>>(goal_cells, n, G, order);
    return cudaGetLastError();
}

}  // namespace ffmp
