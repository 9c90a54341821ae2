// scenario.cu — SPEC.md §3: hash-RNG obstacle map + start/goal sampling (replaces the external
// episode_manager node the reference only talks to over ROS, /root/reference/src/train.py:86-90,128-132).
//
// One CTA per work item.  Thread 0 samples start/goal (a handful of hashes), all threads then emit the
// occupancy plane as coalesced 32-bit stores (4 cells per thread per iteration, one hash per obstacle
// block).  HBM-bound: 1 B/cell written.
#include "ffmp_kernels.cuh"

namespace ffmp {

__global__ void __launch_bounds__(256) scenario_kernel(ScenarioArgs a) {
    __shared__ int s_cells[4];
    __shared__ uint32_t s_key;
    const int count = a.count_ptr ? static_cast<int>(*a.count_ptr) : a.count;
    const int G = a.G;
    const int words_per_row = G >> 2;
    const int words = G * words_per_row;
    for (int item = blockIdx.x; item < count; item += gridDim.x) {
        const uint32_t env = a.env_idx ? a.env_idx[item] : static_cast<uint32_t>(item);
        const uint32_t episode = a.episode ? a.episode[item] : a.episode_const;
        const uint32_t gid = a.env_gid ? a.env_gid[item] : a.env_id_base + env;
        const size_t plane = a.slot_mode ? static_cast<size_t>(episode % a.S) * a.N + env : static_cast<size_t>(item);
        if (threadIdx.x == 0) {
            const uint32_t key = scenario_key(a.seed, gid, episode);
            const ScenarioParams sp = sample_scenario(key, G, a.goal_mode);
            store_scenario_record(a.scen + plane * SC_WORDS, sp, key);
            const int si = sp.si, sj = sp.sj, gi = sp.gi, gj = sp.gj;
            s_cells[0] = si; s_cells[1] = sj; s_cells[2] = gi; s_cells[3] = gj;
            s_key = key;
        }
        __syncthreads();
        const int si = s_cells[0], sj = s_cells[1], gi = s_cells[2], gj = s_cells[3];
        const uint32_t key = s_key;
        uint32_t *out = reinterpret_cast<uint32_t *>(a.occ + plane * static_cast<size_t>(G) * G);
        for (int q = threadIdx.x; q < words; q += blockDim.x) {
            const int i = q / words_per_row;
            const int j0 = (q - i * words_per_row) << 2;
            uint32_t word = 0;
            uint32_t h = 0;
            int hblk = -1;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const int j = j0 + c;
                uint32_t o;
                if (i == 0 || j == 0 || i == G - 1 || j == G - 1) o = 1;
                else if ((abs(i - si) <= 2 && abs(j - sj) <= 2) || (abs(i - gi) <= 2 && abs(j - gj) <= 2)) o = 0;
                else {
                    const int jb = j >> a.block_shift;
                    if (jb != hblk) {
                        const uint32_t blk = (static_cast<uint32_t>(i >> a.block_shift) << 16) | static_cast<uint32_t>(jb);
                        h = mix32(key + blk * 0x9E3779B1u);
                        hblk = jb;
                    }
                    o = h < a.p_thresh ? 1u : 0u;
                }
                word |= o << (8 * c);
            }
            out[q] = word;
        }
        __syncthreads();
    }
}

cudaError_t launch_scenarios(const ScenarioArgs &a, int grid, cudaStream_t st) {
    if (grid <= 0) return cudaSuccess;
    scenario_kernel<<<grid, 256, 0, st>>>(a);
    return cudaGetLastError();
}

}  // namespace ffmp
