// learner_feed.cu — SURVEY.md §8(f) row 1: the observation as the learner's input tensor.
// The reference builds `observe_m` every tick as torch.cat of the last two frames, `.float()`, NCHW (1, 2, 100, 100), and
// copies it to the GPU (/root/reference/src/train.py:474-486, 539-545).  Here the two frames are already adjacent in the
// device frame ring (frames[:, p-1 : p+1]); this kernel widens them to float32 or bfloat16 [N][2][W][W] in one pass:
// coalesced 4-byte loads, 16 (8) byte stores, 1 B read + 4 (2) B written per pixel, HBM-bound.  Values are the integers 0..255 times
// `scale` (1.0 reproduces the reference's plain `.float()`; bf16 holds 0..255 exactly).
#include <cuda_bf16.h>

#include "ffmp_kernels.cuh"

namespace ffmp {

namespace {

// Lane l of a warp converts 4 pixels (one 32-bit load) per step, so a warp reads 128 contiguous bytes and writes 512 (f32) /
// 256 (bf16) contiguous bytes per instruction: every store fills whole 32-byte sectors (16-byte stores at a 64-byte stride
// would half-fill them: measured 3.5 TB/s instead of ~6).
template <bool BF16>
__global__ void __launch_bounds__(256) learner_input_kernel(FeedArgs a) {
    constexpr int UN = 4;                                                      // loads in flight per thread
    const int per_env = 2 * a.W * a.W / 4;                                     // 4-pixel words of one env's two frames
    const size_t ring_stride = static_cast<size_t>(a.K) * a.W * a.W;           // bytes between envs in the frame ring
    const size_t first = static_cast<size_t>(a.slot_new - 1) * a.W * a.W;      // the older of the two frames
    for (int n = blockIdx.x; n < a.N; n += gridDim.x) {                        // one env per CTA pass: no index division
        const uint32_t *src = reinterpret_cast<const uint32_t *>(a.frames + n * ring_stride + first);
        const size_t obase = static_cast<size_t>(n) * per_env;
        for (int k0 = threadIdx.x; k0 < per_env; k0 += UN * 256) {
            uint32_t x[UN];
#pragma unroll
            for (int i = 0; i < UN; ++i)
                if (k0 + i * 256 < per_env) x[i] = __ldcs(src + k0 + i * 256);
#pragma unroll
            for (int i = 0; i < UN; ++i) {
                const int k = k0 + i * 256;
                if (k >= per_env) break;
                float f[4];
#pragma unroll
                for (int b = 0; b < 4; ++b) {
                    // byte -> float without I2F: place the byte in the mantissa of 2^23 (PRMT) and subtract 2^23 (exact)
                    const float v = __fsub_rn(__uint_as_float(__byte_perm(x[i], 0x4B000000u, 0x7440 + b)), 8388608.0f);
                    f[b] = __fmul_rn(v, a.scale);
                }
                if (BF16) {
                    const __nv_bfloat162 p0 = __floats2bfloat162_rn(f[0], f[1]), p1 = __floats2bfloat162_rn(f[2], f[3]);
                    __stcs(reinterpret_cast<uint2 *>(static_cast<uint16_t *>(a.out) + (obase + k) * 4),
                           make_uint2(*reinterpret_cast<const uint32_t *>(&p0), *reinterpret_cast<const uint32_t *>(&p1)));
                } else {
                    __stcs(reinterpret_cast<float4 *>(static_cast<float *>(a.out) + (obase + k) * 4), make_float4(f[0], f[1], f[2], f[3]));
                }
            }
        }
    }
}

}  // namespace

cudaError_t launch_learner_input(const FeedArgs &a, cudaStream_t st) {
    if (a.N <= 0) return cudaSuccess;
    const int grid = 148 * 16;
    if (a.bf16) learner_input_kernel<true><<<grid, 256, 0, st>>>(a);
    else learner_input_kernel<false><<<grid, 256, 0, st>>>(a);
    return cudaGetLastError();
}

}  // namespace ffmp
