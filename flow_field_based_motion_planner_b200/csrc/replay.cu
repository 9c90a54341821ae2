// replay.cu — minibatch gather of the device replay ring (the reference's ReplayMemory.sample +
// Brain.make_minibatch, /root/reference/src/train.py:224-225, 349-369): ONE kernel reads the sampled transitions out of the
// ring of packed transition blocks (feed.cu layout) and writes them in the learner's formats — the two observation stacks as
// bf16 NCHW [B][2][W][W] (what the Q network's first convolution consumes; the u8 -> bf16 widening of train.py:544-545 is
// fused into the gather), goal / velocity / reward as f32, done as u8, action as i64.
#include <cuda_bf16.h>

#include "../../include/ffmp_b200.h"
#include "ffmp_kernels.cuh"

namespace ffmp {
namespace {

struct GatherArgs {
    const uint8_t *blocks;       // [T][stride]: frames [N][2][W][W] | rel_goal f32 [N][2] | velocity f32 [N][2] | reward f32 [N] | done u8 [N]
    const int64_t *actions;      // [T][N]
    const int64_t *index;        // [B][2] = (push number k >= 1, env e): state = push k-1, action / reward / done / next = push k
    size_t stride;
    int T, N, W, B;
    __nv_bfloat16 *state_m, *observe_m;
    float *state_g, *state_v, *observe_g, *observe_v, *reward;
    uint8_t *done;
    int64_t *action;
};

__global__ void __launch_bounds__(256) replay_gather_kernel(GatherArgs a) {
    const int b = blockIdx.x;
    const long long k = a.index[2 * b];
    const int e = static_cast<int>(a.index[2 * b + 1]);
    const int t1 = static_cast<int>(k % a.T), t0 = static_cast<int>((k - 1) % a.T);
    const size_t ww2 = static_cast<size_t>(2) * a.W * a.W;
    const uint8_t *b0 = a.blocks + static_cast<size_t>(t0) * a.stride, *b1 = a.blocks + static_cast<size_t>(t1) * a.stride;
    const uchar4 *s0 = reinterpret_cast<const uchar4 *>(b0 + static_cast<size_t>(e) * ww2);
    const uchar4 *s1 = reinterpret_cast<const uchar4 *>(b1 + static_cast<size_t>(e) * ww2);
    uint2 *d0 = reinterpret_cast<uint2 *>(a.state_m + static_cast<size_t>(b) * ww2);
    uint2 *d1 = reinterpret_cast<uint2 *>(a.observe_m + static_cast<size_t>(b) * ww2);
    auto widen = [](uchar4 v) {
        const __nv_bfloat162 lo = __floats2bfloat162_rn(static_cast<float>(v.x), static_cast<float>(v.y));
        const __nv_bfloat162 hi = __floats2bfloat162_rn(static_cast<float>(v.z), static_cast<float>(v.w));
        return make_uint2(*reinterpret_cast<const uint32_t *>(&lo), *reinterpret_cast<const uint32_t *>(&hi));
    };
    for (int i = threadIdx.x; i < static_cast<int>(ww2 / 4); i += blockDim.x) {
        d0[i] = widen(__ldg(s0 + i));
        d1[i] = widen(__ldg(s1 + i));
    }
    if (threadIdx.x == 0) {
        const size_t off = static_cast<size_t>(a.N) * ww2;
        auto f2 = [&](const uint8_t *blk, size_t o) { return *reinterpret_cast<const float2 *>(blk + off + o + static_cast<size_t>(e) * 8); };
        *reinterpret_cast<float2 *>(a.state_g + 2 * b) = f2(b0, 0);
        *reinterpret_cast<float2 *>(a.state_v + 2 * b) = f2(b0, static_cast<size_t>(8) * a.N);
        *reinterpret_cast<float2 *>(a.observe_g + 2 * b) = f2(b1, 0);
        *reinterpret_cast<float2 *>(a.observe_v + 2 * b) = f2(b1, static_cast<size_t>(8) * a.N);
        a.reward[b] = *reinterpret_cast<const float *>(b1 + off + static_cast<size_t>(16) * a.N + static_cast<size_t>(e) * 4);
        a.done[b] = b1[off + static_cast<size_t>(20) * a.N + e];
        a.action[b] = a.actions[static_cast<size_t>(t1) * a.N + e];
    }
}

}  // namespace
}  // namespace ffmp

extern "C" int ffmp_replay_gather(int32_t device, const uint8_t *blocks_dev, size_t stride, int32_t T, int32_t N, int32_t W,
                                  const int64_t *actions_dev, const int64_t *index_dev, int32_t B, void *state_m_bf16,
                                  void *observe_m_bf16, float *state_g, float *state_v, float *observe_g, float *observe_v,
                                  float *reward, uint8_t *done, int64_t *action, void *stream) {
    if (!blocks_dev || !actions_dev || !index_dev || !state_m_bf16 || !observe_m_bf16 || !state_g || !state_v || !observe_g ||
        !observe_v || !reward || !done || !action || B < 0 || T < 2 || N <= 0 || W <= 0 || (W % 2) || (stride % 16))
        return -1;       // FFMP_ERR_ARG
    if (B == 0) return 0;
    int prev = 0;
    cudaGetDevice(&prev);
    cudaSetDevice(device);
    ffmp::GatherArgs a{blocks_dev, actions_dev, index_dev, stride, T, N, W, B, static_cast<__nv_bfloat16 *>(state_m_bf16),
                       static_cast<__nv_bfloat16 *>(observe_m_bf16), state_g, state_v, observe_g, observe_v, reward, done, action};
    ffmp::replay_gather_kernel<<<B, 256, 0, static_cast<cudaStream_t>(stream)>>>(a);
    const cudaError_t ce = cudaGetLastError();
    cudaSetDevice(prev);
    return ce == cudaSuccess ? 0 : -3;      // FFMP_ERR_CUDA
}
