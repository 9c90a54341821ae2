// feed.cu — SURVEY.md §8(e): the learner feed as ONE kernel over NVLink peer memory.
//
// What it replaces: pack_transitions (five strided torch copies into a staging tensor) followed by an NCCL
// all_gather_into_tensor (sharding.py; measured 0.40 ms per step for 82 MB per rank at 2 GPUs).  Here every rank owns a
// gather buffer [world][block] that its peers map through CUDA IPC; feed_push_kernel reads the rank's transition block ONCE
// from where the step kernel left it (the two newest frames of the observation ring, relative_goal, velocity, reward, done) and
// stores it, already in the packed layout [maps u8 N*2*W*W | rel_goal f32 N*2 | velocity f32 N*2 | reward f32 N | done u8 N],
// straight into slot `rank` of every destination's buffer with 16-byte stores over NVLink (or locally for its own copy).  The
// last CTA then publishes the push sequence number in every destination's flag word (fence at system scope first).
// Flow control is credit based and collective-free: two buffers alternate by sequence parity, a consumer releases a buffer
// by writing the sequence number it has finished reading into every producer's ack word, and a producer waits (a one-warp
// spin kernel on its stream, bounded by a timeout) until the buffer it is about to overwrite has been released.
#include <cstdlib>

#include "ffmp_kernels.cuh"

namespace ffmp {

namespace {

__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t *p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// One warp; lane r < world watches word r.  Returns when every watched word has reached `want` (sequence numbers only grow;
// the comparison is wrap-safe) or after `timeout_ns`, in which case bit 1 of the error word is set.
__global__ void __launch_bounds__(32) feed_spin_kernel(const uint32_t *words, int stride_words, uint32_t mask, uint32_t want,
                                                       unsigned long long timeout_ns, uint32_t *error_word) {
    const int lane = threadIdx.x;
    const bool watch = (mask >> lane) & 1u;
    unsigned long long t0;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    for (;;) {
        bool ok = true;
        if (watch) ok = static_cast<int32_t>(ld_acquire_sys(words + static_cast<size_t>(lane) * stride_words) - want) >= 0;
        if (__all_sync(FULL, ok)) return;
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        if (t - t0 > timeout_ns) {
            if (lane == 0) atomicOr(error_word, 2u);
            return;
        }
        __nanosleep(200);
    }
}

// lane r < world with bit r of mask set stores `value` to targets[r] (a word in rank r's memory)
__global__ void __launch_bounds__(32) feed_signal_kernel(FeedTargets t, uint32_t mask, uint32_t value) {
    const int lane = threadIdx.x;
    __threadfence_system();
    if (lane < FEED_MAX_WORLD && ((mask >> lane) & 1u))
        asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(t.word[lane]), "r"(value) : "memory");
}

__global__ void __launch_bounds__(256) feed_push_kernel(FeedPushArgs a) {
    const int ndst = a.ndst;
    // ---- maps: env e contributes 2 W^2 contiguous bytes (frames[e][p-1 .. p]) -> offset e * 2 W^2 of the block ----
    const int per_env16 = (2 * a.W * a.W) >> 4;
    const size_t ring_stride = static_cast<size_t>(a.K) * a.W * a.W;
    const size_t first = static_cast<size_t>(a.slot_new - 1) * a.W * a.W;
    for (int e = blockIdx.x; e < a.N; e += gridDim.x) {
        const uint4 *src = reinterpret_cast<const uint4 *>(a.frames + e * ring_stride + first);
        const size_t off16 = static_cast<size_t>(e) * per_env16;
        for (int k0 = threadIdx.x; k0 < per_env16; k0 += 4 * 256) {
            uint4 x[4];
#pragma unroll
            for (int i = 0; i < 4; ++i)
                if (k0 + i * 256 < per_env16) x[i] = __ldcs(src + k0 + i * 256);
#pragma unroll
            for (int i = 0; i < 4; ++i)
                if (k0 + i * 256 < per_env16)
                    for (int d = 0; d < ndst; ++d) reinterpret_cast<uint4 *>(a.dst[d])[off16 + k0 + i * 256] = x[i];
        }
    }
    // ---- the small per-env fields: 21 N bytes behind the maps, contiguous in the block and in the env's output block ----
    {
        const size_t maps_bytes = static_cast<size_t>(a.N) * 2 * a.W * a.W;
        const int gtid = blockIdx.x * blockDim.x + threadIdx.x, gsz = gridDim.x * blockDim.x;
        const int N = a.N;
        // rel_goal | velocity (8 N + 8 N bytes), reward (4 N), done (N): 32-bit words where possible
        for (int i = gtid; i < 2 * N; i += gsz) {
            const uint32_t g = reinterpret_cast<const uint32_t *>(a.rel_goal)[i], v = reinterpret_cast<const uint32_t *>(a.velocity)[i];
            for (int d = 0; d < ndst; ++d) {
                uint32_t *o = reinterpret_cast<uint32_t *>(a.dst[d] + maps_bytes);
                o[i] = g; o[2 * N + i] = v;
            }
        }
        for (int i = gtid; i < N; i += gsz) {
            const uint32_t r = reinterpret_cast<const uint32_t *>(a.reward)[i];
            const uint8_t dn = a.done[i];
            for (int d = 0; d < ndst; ++d) {
                reinterpret_cast<uint32_t *>(a.dst[d] + maps_bytes)[4 * N + i] = r;
                (a.dst[d] + maps_bytes + 20 * static_cast<size_t>(N))[i] = dn;
            }
        }
    }
    if (!a.ticket) return;            // local packing only (ffmp_pack_transitions): nothing to publish
    // ---- publish: the last CTA to finish writes the sequence number into every destination's flag word ----
    __threadfence_system();
    __syncthreads();
    __shared__ bool last;
    if (threadIdx.x == 0) last = atomicInc(a.ticket, gridDim.x - 1) == gridDim.x - 1;
    __syncthreads();
    if (last && static_cast<int>(threadIdx.x) < ndst) {
        __threadfence_system();
        asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(a.flag[threadIdx.x]), "r"(a.seq) : "memory");
    }
}

}  // namespace

cudaError_t launch_feed_spin(const uint32_t *words, int stride_words, uint32_t mask, uint32_t want, double timeout_s,
                             uint32_t *error_word, cudaStream_t st) {
    feed_spin_kernel<<<1, 32, 0, st>>>(words, stride_words, mask, want, static_cast<unsigned long long>(timeout_s * 1e9), error_word);
    return cudaGetLastError();
}

cudaError_t launch_feed_signal(const FeedTargets &t, uint32_t mask, uint32_t value, cudaStream_t st) {
    feed_signal_kernel<<<1, 32, 0, st>>>(t, mask, value);
    return cudaGetLastError();
}

cudaError_t launch_feed_push(const FeedPushArgs &a, cudaStream_t st) {
    if (a.N <= 0 || a.ndst <= 0) return cudaSuccess;
    static int per_sm = 0;
    if (per_sm == 0) {
        const char *e = std::getenv("FFMP_FEED_CTAS_PER_SM");      // tuning knob (tools/feed_bench.py)
        per_sm = e && std::atoi(e) > 0 ? std::atoi(e) : 4;
    }
    const int grid = a.N < 148 * per_sm ? a.N : 148 * per_sm;
    feed_push_kernel<<<grid, 256, 0, st>>>(a);
    return cudaGetLastError();
}

}  // namespace ffmp
