// qnet.cu — SURVEY.md §8(f) row 2: the policy / Q network of the reference trainer, `Network.forward`
// (/root/reference/src/train.py:231-303: conv 32x32 -> conv 32x32 -> conv 8x8 -> [+ scalar tile] -> conv 8x8 applied three
// times -> fc 6400-512-512 -> dueling heads 28 + 1), as hand-written sm_100a kernels.
//
// Every layer is ONE kernel, `qnet_gemm_kernel`: an implicit GEMM  D[M, N] = act(A[M, K] * B[N, K]^T + bias)  on the 5th
// generation tensor cores.
//   * A (activations, bf16, NHWC) is never unfolded in memory.  A K-slice of 64 elements of an output pixel's receptive field
//     is 128 contiguous bytes of the NHWC tensor ((kw, c) run along an input row), so the "im2col matrix" of a tile of output
//     pixels is one 4-D TMA box: d0 = position inside the input row's (kw, c) run, d1 = output column (stride = one pixel:
//     the dimension OVERLAPS d0 — tools/tmap_probe.cu checks that the driver and the copy engine accept it), d2 = input row,
//     d3 = sample.  The box lands in shared memory in the 128-byte-swizzled K-major layout tcgen05.mma reads.
//   * B (weights, bf16, [N][K] with K ordered (kh, kw, c)) is a plain 2-D TMA box.
//   * warp 0 = TMA producer, warp 1 = MMA issuer (one elected thread: tcgen05.mma, M = 128, N = 32 / 64, K = 16, fp32
//     accumulators in TMEM, tcgen05.commit onto the stage's mbarrier), warps 2..5 = epilogue (tcgen05.ld, bias, ReLU, the
//     scalar tile of train.py:264-267, bf16 / fp32 store).  A 6-stage shared-memory ring decouples the three.
// The fully connected layers are the same kernel with a rank-2 activation "image".
#include <cuda.h>
#include <cuda_bf16.h>

#include <cstdio>
#include <cstdlib>
#include <new>
#include <string>

#include "../../include/ffmp_b200.h"
#include "flow_bits.cuh"

namespace ffmp {
namespace {

constexpr int QG_STAGES = 6;
constexpr int QG_A_BYTES = 128 * 128;          // 128 rows x 64 bf16
constexpr int QG_THREADS = 192;

struct QGemmArgs {
    int k_outer, k_inner;        // K loop: k_outer kernel rows x k_inner 64-element slices per row
    int a_c1_step, a_c2_step;    // A box origin of tile t: (64 * ki, t * a_c1_step, t * a_c2_step + ko, blockIdx.y)
    int box1, box2;              // rows of the A box = box1 * box2 (<= 128): row r <-> (i1 = r % box1, i2 = r / box1)
    int lim1, lim2;              // valid extents of (t * a_c1_step + i1, t * a_c2_step + i2)
    int ldc;                     // output row pitch in elements; output row = (blockIdx.y * lim2 + idx2) * lim1 + idx1
    int relu, out_f32;
    const float *bias;           // [N]
    const float *post_add;       // device scalar added after the activation (train.py:264-276) or null
    void *out;
    int dbg_shift, dbg_bo;       // experiment (QNET_DBG_SHIFT / QNET_DBG_BO): A rows shifted in shared memory, descriptor start unaligned
};

__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap *tm, int c0, int c1, int c2, int c3, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
                 ::"r"(dst), "l"(tm), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(bar) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap *tm, int c0, int c1, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(dst), "l"(tm), "r"(c0), "r"(c1), "r"(bar) : "memory");
}
// K-major, 128-byte swizzle: rows of 128 bytes, 8-row groups 1024 bytes apart (SBO), descriptor version 1 (sm_100)
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) {
    return static_cast<uint64_t>((smem_addr >> 4) & 0x3FFFu) | (static_cast<uint64_t>(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// the same with the descriptors as (lo, hi) halves: an issue loop advances `lo` with one 32-bit add per MMA
__device__ __forceinline__ void umma_f16_lh(uint32_t tmem_d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi, uint32_t idesc,
                                            uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\tsetp.ne.b32 p, %6, 0;\n\tmov.b64 da, {%1, %2};\n\tmov.b64 db, {%3, %4};\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t}"
                 ::"r"(tmem_d), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
                   "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

template <int BN>
__global__ void __launch_bounds__(QG_THREADS, 1)
qnet_gemm_kernel(const __grid_constant__ CUtensorMap tma, const __grid_constant__ CUtensorMap tmb, QGemmArgs g) {
    extern __shared__ __align__(1024) uint8_t smem[];
    constexpr int B_BYTES = BN * 128;
    uint8_t *sa = smem, *sb = smem + QG_STAGES * QG_A_BYTES;
    uint64_t *bars = reinterpret_cast<uint64_t *>(sb + QG_STAGES * B_BYTES);     // full[ST], empty[ST], accum
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 2 * QG_STAGES + 1);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    auto bar_addr = [&](int i) { return static_cast<uint32_t>(__cvta_generic_to_shared(bars + i)); };
    const uint32_t sa_s = static_cast<uint32_t>(__cvta_generic_to_shared(sa));
    const uint32_t sb_s = static_cast<uint32_t>(__cvta_generic_to_shared(sb));

    if (threadIdx.x == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tma) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmb) : "memory");
        for (int i = 0; i < 2 * QG_STAGES + 1; ++i) mbar_init(bar_addr(i), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {          // TMEM: BN fp32 accumulator columns (a power of two >= 32), allocated and freed by this warp
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                     ::"r"(static_cast<uint32_t>(__cvta_generic_to_shared(tmem_slot))), "r"(static_cast<uint32_t>(BN)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_d = *tmem_slot;

    const int nsteps = g.k_outer * g.k_inner;
    const int tile = blockIdx.x, n0 = blockIdx.z * BN;
    if (warp == 0) {
        if (lane == 0) {
            const uint32_t tx = static_cast<uint32_t>(g.box1 * g.box2 * 128 + B_BYTES);
            int ko = 0, ki = 0;
            for (int s = 0; s < nsteps; ++s) {
                const int st = s % QG_STAGES, use = s / QG_STAGES;
                if (use > 0) mbar_wait(bar_addr(QG_STAGES + st), static_cast<uint32_t>((use - 1) & 1));
                mbar_expect_tx(bar_addr(st), tx);
                tma_load_4d(sa_s + st * QG_A_BYTES, &tma, 64 * ki, tile * g.a_c1_step - g.dbg_shift, tile * g.a_c2_step + ko,
                            static_cast<int>(blockIdx.y), bar_addr(st));
                tma_load_2d(sb_s + st * B_BYTES, &tmb, 64 * s, n0, bar_addr(st));
                if (++ki == g.k_inner) { ki = 0; ++ko; }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            // instruction descriptor: D fp32, A / B bf16, both K-major, N = BN, M = 128
            constexpr uint32_t IDESC = (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(BN >> 3) << 17) | (8u << 24);
            for (int s = 0; s < nsteps; ++s) {
                const int st = s % QG_STAGES, use = s / QG_STAGES;
                mbar_wait(bar_addr(st), static_cast<uint32_t>(use & 1));
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t a_addr = sa_s + st * QG_A_BYTES + g.dbg_shift * 128;
                const uint64_t ad = umma_desc_sw128(a_addr) | (g.dbg_bo ? static_cast<uint64_t>((a_addr >> 7) & 7u) << 49 : 0ull);
                const uint64_t bd = umma_desc_sw128(sb_s + st * B_BYTES);
#pragma unroll
                for (int k = 0; k < 4; ++k)       // 4 x (K = 16): +32 bytes inside the 128-byte swizzle atom
                    umma_f16(tmem_d, ad + 2 * k, bd + 2 * k, IDESC, (s | k) ? 1u : 0u);
                umma_commit(bar_addr(QG_STAGES + st));       // the stage is free once these MMAs have read it
            }
            umma_commit(bar_addr(2 * QG_STAGES));            // accumulator complete
        }
    } else {
        // epilogue: warp w may touch TMEM lanes 32 * (w % 4) .. + 31 only; thread <-> accumulator row
        const int q = warp & 3, r = 32 * q + lane;
        mbar_wait(bar_addr(2 * QG_STAGES), 0u);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const int i1 = r % g.box1, i2 = r / g.box1;
        const int idx1 = tile * g.a_c1_step + i1, idx2 = tile * g.a_c2_step + i2;
        const bool valid = r < g.box1 * g.box2 && idx1 < g.lim1 && idx2 < g.lim2;
        const size_t orow = (static_cast<size_t>(blockIdx.y) * g.lim2 + idx2) * g.lim1 + idx1;
        const float add = g.post_add ? *g.post_add : 0.0f;
#pragma unroll
        for (int c = 0; c < BN / 16; ++c) {
            uint32_t v[16];
            tmem_ld16(tmem_d + (static_cast<uint32_t>(32 * q) << 16) + 16 * c, v);
            float f[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                float x = __uint_as_float(v[j]) + g.bias[n0 + 16 * c + j];
                if (g.relu) x = fmaxf(x, 0.0f);
                f[j] = x + add;
            }
            if (valid) {
                if (g.out_f32) {
                    float4 *dst = reinterpret_cast<float4 *>(static_cast<float *>(g.out) + orow * g.ldc + n0 + 16 * c);
#pragma unroll
                    for (int j = 0; j < 4; ++j) dst[j] = make_float4(f[4 * j], f[4 * j + 1], f[4 * j + 2], f[4 * j + 3]);
                } else {
                    uint32_t p[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const __nv_bfloat162 h = __floats2bfloat162_rn(f[2 * j], f[2 * j + 1]);
                        p[j] = *reinterpret_cast<const uint32_t *>(&h);
                    }
                    uint4 *dst = reinterpret_cast<uint4 *>(static_cast<__nv_bfloat16 *>(g.out) + orow * g.ldc + n0 + 16 * c);
                    dst[0] = make_uint4(p[0], p[1], p[2], p[3]);
                    dst[1] = make_uint4(p[4], p[5], p[6], p[7]);
                }
            }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    }
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"(static_cast<uint32_t>(BN)) : "memory");
    }
}

// ---- conv1 (2 -> 32 channels, 32 x 32 kernel, train.py:236) with the whole input image resident in shared memory ------------
// The generic kernel re-fetches every input pixel 32 x 32 times through L2 (one 128-byte K-slice per output pixel and kernel
// tap pair): 29 GB per 256 samples, L2-bound at 4x the tensor time.  Here one CTA owns one sample: the NHWC8 image (100 x 100
// pixels of 16 bytes = 160 KB) is copied to shared memory ONCE, in its global layout, and every A operand is a VIEW of it.
// With the un-swizzled K-major layout a tcgen05 operand is made of 8-row x 16-byte core matrices whose rows are 16 bytes apart:
// exactly consecutive pixels.  M row m of a tile <-> image pixel m0 + m (row pitch 100: the 31 rightmost of every 100 rows
// are windows that run over the image edge and are discarded), K = 16 <-> the two taps kw, kw + 1, i.e. the pixels
// m0 + m + kh * 100 + kw (+1): descriptor start = that pixel, leading byte offset 16 (next pixel), stride byte offset 128 (next
// 8 rows).  Overlapping "rows" are only address arithmetic to the tensor core.  The weights of one kernel row (32 taps x 8
// channels x 32 output channels = 16 KB, pre-arranged in the core-matrix order) stream through a 2-stage ring; 16 accumulators
// of 32 columns fill TMEM, so a pass covers 2048 image pixels and a sample takes 4 passes of 32 x 16 x 16 MMAs.
constexpr int C1_IMG = 100 * 100 * 16;            // bytes of the NHWC8 image
constexpr int C1_IMG_PAD = C1_IMG + 8192;         // windows of discarded rows may read past the image
constexpr int C1_WROW = 32 * 32 * 16;             // weights of one kernel row: [kw 32][oc 32][8 ch] bf16
constexpr int C1_TILES = (68 * 100 + 68) / 128 + 1;   // 54 tiles of 128 pixels cover the last valid output pixel
constexpr int C1_ACC = 16;                        // accumulators (32 TMEM columns each) per pass
constexpr int C1_SMEM = C1_IMG_PAD + 2 * C1_WROW + 128;

__device__ __forceinline__ uint64_t umma_desc_nosw(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return static_cast<uint64_t>((smem_addr >> 4) & 0x3FFFu) | (static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFFu) << 16) |
           (static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFFu) << 32) | (1ull << 46);
}

// planar = 1: output as [4 channel chunks][69 * 69 pixels][8 channels] (the layout conv2's resident kernel views), else NHWC
__global__ void __launch_bounds__(QG_THREADS, 1)
qnet_conv1_kernel(const __nv_bfloat16 *in, const __nv_bfloat16 *w, const float *bias, __nv_bfloat16 *out, int planar) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t *img = smem, *wst = smem + C1_IMG_PAD;
    uint64_t *bars = reinterpret_cast<uint64_t *>(wst + 2 * C1_WROW);     // img, wfull[2], wempty[2], accfull, accempty
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 8);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    auto bar_addr = [&](int i) { return static_cast<uint32_t>(__cvta_generic_to_shared(bars + i)); };
    const uint32_t img_s = static_cast<uint32_t>(__cvta_generic_to_shared(img)), wst_s = static_cast<uint32_t>(__cvta_generic_to_shared(wst));
    const int n = blockIdx.x;
    if (threadIdx.x == 0) {
        for (int i = 0; i < 6; ++i) mbar_init(bar_addr(i), 1);
        mbar_init(bar_addr(6), 4);                // accempty: one arrival per epilogue warp
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                     ::"r"(static_cast<uint32_t>(__cvta_generic_to_shared(tmem_slot))), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    // the pad behind the image is read by discarded rows only, but must not hold NaN patterns that trap nothing — just define it
    for (int i = threadIdx.x; i < (C1_IMG_PAD - C1_IMG) / 16; i += QG_THREADS) reinterpret_cast<uint4 *>(img + C1_IMG)[i] = make_uint4(0, 0, 0, 0);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_d = *tmem_slot;
    constexpr int PASSES = (C1_TILES + C1_ACC - 1) / C1_ACC;

    if (warp == 0) {
        if (lane == 0) {
            mbar_expect_tx(bar_addr(0), C1_IMG);
            const uint8_t *src = reinterpret_cast<const uint8_t *>(in) + static_cast<size_t>(n) * C1_IMG;
            for (int c = 0; c < 10; ++c) tma_bulk_g2s(img_s + c * 16000, src + c * 16000, 16000, bar_addr(0));
            int s = 0;
            for (int p = 0; p < PASSES; ++p)
                for (int kh = 0; kh < 32; ++kh, ++s) {
                    const int st = s & 1, use = s >> 1;
                    if (use > 0) mbar_wait(bar_addr(3 + st), static_cast<uint32_t>((use - 1) & 1));
                    mbar_expect_tx(bar_addr(1 + st), C1_WROW);
                    tma_bulk_g2s(wst_s + st * C1_WROW, reinterpret_cast<const uint8_t *>(w) + static_cast<size_t>(kh) * C1_WROW, C1_WROW, bar_addr(1 + st));
                }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t IDESC = (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(32 >> 3) << 17) | (8u << 24);
            mbar_wait(bar_addr(0), 0u);
            int s = 0;
            for (int p = 0; p < PASSES; ++p) {
                if (p > 0) mbar_wait(bar_addr(6), static_cast<uint32_t>((p - 1) & 1));     // the epilogue drained the accumulators
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const int nt = min(C1_ACC, C1_TILES - p * C1_ACC);
                for (int kh = 0; kh < 32; ++kh, ++s) {
                    const int st = s & 1, use = s >> 1;
                    mbar_wait(bar_addr(1 + st), static_cast<uint32_t>(use & 1));
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    // descriptor halves: hi = (stride byte offset, version), lo = (start >> 4, leading byte offset); only the start moves
                    const uint64_t bd0 = umma_desc_nosw(wst_s + st * C1_WROW, 512, 128);
                    const uint64_t ad0 = umma_desc_nosw(img_s + static_cast<uint32_t>((p * C1_ACC * 128 + kh * 100) * 16), 16, 128);
                    const uint32_t b_hi = static_cast<uint32_t>(bd0 >> 32), a_hi = static_cast<uint32_t>(ad0 >> 32);
                    uint32_t b_lo = static_cast<uint32_t>(bd0), a_lo0 = static_cast<uint32_t>(ad0);
#pragma unroll 1
                    for (int kp = 0; kp < 16; ++kp) {
                        uint32_t a_lo = a_lo0;
                        const uint32_t acc = (kh | kp) ? 1u : 0u;
#pragma unroll 4
                        for (int t = 0; t < nt; ++t) {
                            umma_f16_lh(tmem_d + 32 * t, a_lo, a_hi, b_lo, b_hi, IDESC, acc);
                            a_lo += 2048 >> 4;                 // next tile: 128 pixels
                        }
                        a_lo0 += 32 >> 4;                      // next tap pair: 2 pixels
                        b_lo += 1024 >> 4;                     // next two 512-byte tap chunks
                    }
                    umma_commit(bar_addr(3 + st));
                }
                umma_commit(bar_addr(5));
            }
        }
    } else {
        const int q = warp & 3;
        for (int p = 0; p < PASSES; ++p) {
            mbar_wait(bar_addr(5), static_cast<uint32_t>(p & 1));
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const int nt = min(C1_ACC, C1_TILES - p * C1_ACC);
            for (int t = 0; t < nt; ++t) {
                const int m = (p * C1_ACC + t) * 128 + 32 * q + lane;
                const int oh = m / 100, ow = m - oh * 100;
                const bool valid = oh < 69 && ow < 69;
                uint32_t pk[16];
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    uint32_t v[16];
                    tmem_ld16(tmem_d + (static_cast<uint32_t>(32 * q) << 16) + 32 * t + 16 * c, v);
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const float x0 = fmaxf(__uint_as_float(v[2 * j]) + bias[16 * c + 2 * j], 0.0f);
                        const float x1 = fmaxf(__uint_as_float(v[2 * j + 1]) + bias[16 * c + 2 * j + 1], 0.0f);
                        const __nv_bfloat162 h = __floats2bfloat162_rn(x0, x1);
                        pk[8 * c + j] = *reinterpret_cast<const uint32_t *>(&h);
                    }
                }
                if (valid) {
                    const size_t px = static_cast<size_t>(oh) * 69 + ow;
                    if (planar) {
#pragma unroll
                        for (int c4 = 0; c4 < 4; ++c4)
                            *reinterpret_cast<uint4 *>(out + ((static_cast<size_t>(n) * 4 + c4) * 4761 + px) * 8) =
                                make_uint4(pk[4 * c4], pk[4 * c4 + 1], pk[4 * c4 + 2], pk[4 * c4 + 3]);
                    } else {
                        uint4 *dst = reinterpret_cast<uint4 *>(out + (static_cast<size_t>(n) * 4761 + px) * 32);
#pragma unroll
                        for (int c4 = 0; c4 < 4; ++c4) dst[c4] = make_uint4(pk[4 * c4], pk[4 * c4 + 1], pk[4 * c4 + 2], pk[4 * c4 + 3]);
                    }
                }
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar_addr(6)) : "memory");
        }
    }
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"(512u) : "memory");
    }
}

// conv1 weight [32][2][32][32] f32 -> [kh][kw][oc][8] bf16 (channels 2..7 zero): the core-matrix order of the resident kernel
__global__ void qnet_conv1_weight_core_kernel(const float *w, __nv_bfloat16 *out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= 32 * 32 * 32 * 8) return;
    const int c = i & 7, oc = (i >> 3) & 31, kw = (i >> 8) & 31, kh = i >> 13;
    out[i] = __float2bfloat16(c < 2 ? w[((oc * 2 + c) * 32 + kh) * 32 + kw] : 0.0f);
}

// ---- small helpers around the GEMMs -----------------------------------------------------------------------------------
// state_m bf16 / f32 [B][2][H][W] (NCHW, the layout of learner_input / train.py:544-545) -> NHWC with 8 channels (6 zero)
template <typename T>
__global__ void qnet_pack_input_kernel(const T *in, __nv_bfloat16 *out, int B, int HW) {
    const size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
    if (i >= static_cast<size_t>(B) * HW) return;
    const size_t b = i / HW, p = i - b * HW;
    const float c0 = static_cast<float>(in[(2 * b) * HW + p]), c1 = static_cast<float>(in[(2 * b + 1) * HW + p]);
    const __nv_bfloat162 h = __floats2bfloat162_rn(c0, c1);
    *reinterpret_cast<uint4 *>(out + i * 8) = make_uint4(*reinterpret_cast<const uint32_t *>(&h), 0u, 0u, 0u);
}

// conv weight [OC][IC][KH][KW] f32 -> [OC][KH][KW][ICP] bf16 (zero for ic >= IC)
__global__ void qnet_conv_weight_kernel(const float *w, __nv_bfloat16 *out, int OC, int IC, int KH, int KW, int ICP) {
    const size_t n = static_cast<size_t>(OC) * KH * KW * ICP;
    for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n; i += static_cast<size_t>(gridDim.x) * blockDim.x) {
        const int ic = static_cast<int>(i % ICP);
        size_t t = i / ICP;
        const int kw = static_cast<int>(t % KW); t /= KW;
        const int kh = static_cast<int>(t % KH);
        const int oc = static_cast<int>(t / KH);
        out[i] = __float2bfloat16(ic < IC ? w[((static_cast<size_t>(oc) * IC + ic) * KH + kh) * KW + kw] : 0.0f);
    }
}

// fc2 weight [512][C * HW] with torch.flatten's (c, h, w) column order -> (h, w, c) order of the NHWC activations
__global__ void qnet_fc2_weight_kernel(const float *w, __nv_bfloat16 *out, int N, int C, int HW) {
    const size_t n = static_cast<size_t>(N) * C * HW;
    for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n; i += static_cast<size_t>(gridDim.x) * blockDim.x) {
        const int c = static_cast<int>(i % C);
        const size_t t = i / C;
        const int p = static_cast<int>(t % HW);
        const int o = static_cast<int>(t / HW);
        out[i] = __float2bfloat16(w[(static_cast<size_t>(o) * C + c) * HW + p]);
    }
}

// rows [r0, r0 + rows) of a [.][K] bf16 matrix from an f32 [rows][K] matrix (plain cast); used for fc3 and the two heads
__global__ void qnet_cast_rows_kernel(const float *w, __nv_bfloat16 *out, size_t n) {
    for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n; i += static_cast<size_t>(gridDim.x) * blockDim.x)
        out[i] = __float2bfloat16(w[i]);
}

// train.py:259-267: x_gvt_ = relu(fc1(cat(g, v, t))); only x_gvt_[0][30] (sample 0, unit 30) survives as a scalar tile
__global__ void qnet_scalar_tile_kernel(const float *fc1_w, const float *fc1_b, const float *g, const float *v, const float *t,
                                        int enabled, float *out) {
    if (threadIdx.x) return;
    float s = 0.0f;
    if (enabled) {
        const float *w = fc1_w + 30 * 5;
        s = fc1_b[30] + w[0] * g[0] + w[1] * g[1] + w[2] * v[0] + w[3] * v[1] + w[4] * t[0];
        s = fmaxf(s, 0.0f);
    }
    *out = s;
}

// train.py:286-299: output = adv + val - mean(adv); heads f32 [B][32] = (28 advantages, value, 3 zero columns)
__global__ void qnet_dueling_kernel(const float *heads, float *q, int B) {
    const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (b >= B) return;
    const float x = heads[static_cast<size_t>(b) * 32 + lane];
    float a = lane < 28 ? x : 0.0f;
#pragma unroll
    for (int d = 16; d; d >>= 1) a += __shfl_xor_sync(FULL, a, d);
    const float val = __shfl_sync(FULL, x, 28);
    if (lane < 28) q[static_cast<size_t>(b) * 28 + lane] = x + val - a * (1.0f / 28.0f);
}

typedef CUresult (*EncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                             const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeFn encode_fn() {
    static EncodeFn fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeFn>(p);
    }
    return fn;
}

bool make_map(CUtensorMap *tm, void *base, int rank, const cuuint64_t *dims, const cuuint64_t *strides_bytes, const cuuint32_t *box) {
    const cuuint32_t es[4] = {1, 1, 1, 1};
    EncodeFn fn = encode_fn();
    return fn && fn(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, static_cast<cuuint32_t>(rank), base, dims, strides_bytes, box, es,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

thread_local std::string q_err;
int qfail(int code, const char *what, cudaError_t ce = cudaSuccess) {
    q_err = what;
    if (ce != cudaSuccess) { q_err += ": "; q_err += cudaGetErrorString(ce); }
    return code;
}
#define QCK(call) do { cudaError_t ce_ = (call); if (ce_ != cudaSuccess) return qfail(FFMP_ERR_CUDA, #call, ce_); } while (0)

// geometry of Network.forward on the reference's 100 x 100 maps (train.py:236-239: kernel sizes 32, 32, 8, 8)
struct ConvGeo { int IH, IW, ICP, KH, KW, OC, OH, OW, rows; };     // rows = output rows per tile
constexpr ConvGeo CONVS[6] = {
    {100, 100, 8, 32, 32, 32, 69, 69, 1},     // conv1 (2 input channels padded to 8: 16-byte pixels)
    {69, 69, 32, 32, 32, 64, 38, 38, 3},      // conv2
    {38, 38, 64, 8, 8, 64, 31, 31, 4},        // conv3
    {31, 31, 64, 8, 8, 64, 24, 24, 5},        // conv4, first application
    {24, 24, 64, 8, 8, 64, 17, 17, 7},        // conv4, second
    {17, 17, 64, 8, 8, 64, 10, 10, 12},       // conv4, third
};

}  // namespace
}  // namespace ffmp

using namespace ffmp;

struct ffmp_qnet {
    int device = 0, max_batch = 0;
    bool loaded = false;
    __nv_bfloat16 *w[6] = {nullptr};       // conv1, conv2, conv3, conv4, fc2, fc3 in GEMM layout; heads below
    __nv_bfloat16 *w_heads = nullptr;      // [32][512]: 28 advantage rows, the value row, 3 zero rows
    __nv_bfloat16 *w1core = nullptr;       // conv1 weights in the core-matrix order of qnet_conv1_kernel: [kh][kw][oc][8]
    float *bias[7] = {nullptr};            // conv1..4, fc2, fc3, heads (32)
    float *fc1_w = nullptr, *fc1_b = nullptr, *scalar = nullptr, *heads_out = nullptr;
    __nv_bfloat16 *act[8] = {nullptr};     // NHWC8 input, conv1 .. conv4c outputs, fc2, fc3 outputs
    size_t act_elems[8] = {0};
    uint64_t launches = 0;
};

extern "C" {

const char *ffmp_qnet_last_error(void) { return q_err.c_str(); }

int ffmp_qnet_destroy(ffmp_qnet *n) {
    if (!n) return FFMP_OK;
    int prev = 0;
    cudaGetDevice(&prev);
    cudaSetDevice(n->device);
    for (auto p : n->w) cudaFree(p);
    cudaFree(n->w_heads);
    cudaFree(n->w1core);
    for (auto p : n->bias) cudaFree(p);
    cudaFree(n->fc1_w); cudaFree(n->fc1_b); cudaFree(n->scalar); cudaFree(n->heads_out);
    for (auto p : n->act) cudaFree(p);
    cudaSetDevice(prev);
    delete n;
    return FFMP_OK;
}

int ffmp_qnet_create(int32_t device, int32_t max_batch, ffmp_qnet **out) {
    if (!out || max_batch <= 0) return qfail(FFMP_ERR_ARG, "bad argument");
    *out = nullptr;
    int ndev = 0, major = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) {
        cudaGetLastError();
        return qfail(FFMP_ERR_DEVICE, "no such CUDA device (there is no CPU fallback)");
    }
    cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device);
    if (major != 10) return qfail(FFMP_ERR_DEVICE, "device is not sm_100 (B200); this library carries sm_100a code only");
    if (!encode_fn()) return qfail(FFMP_ERR_CUDA, "cuTensorMapEncodeTiled is not available");
    int prev = 0;
    cudaGetDevice(&prev);
    cudaSetDevice(device);
    ffmp_qnet *n = new (std::nothrow) ffmp_qnet();
    if (!n) return qfail(FFMP_ERR_ARG, "out of host memory");
    n->device = device; n->max_batch = max_batch;
    const size_t B = max_batch;
    const size_t wsz[6] = {32u * 32 * 32 * 8, 64u * 32 * 32 * 32, 64u * 8 * 8 * 64, 64u * 8 * 8 * 64, 512u * 6400, 512u * 512};
    const size_t bsz[7] = {32, 64, 64, 64, 512, 512, 32};
    n->act_elems[0] = B * 100 * 100 * 8;
    for (int l = 0; l < 6; ++l) n->act_elems[1 + l] = B * CONVS[l].OH * CONVS[l].OW * CONVS[l].OC;
    n->act_elems[7] = B * 512 * 2;      // fc2 output, then fc3 output
    cudaError_t ce = cudaSuccess;
    for (int l = 0; l < 6 && ce == cudaSuccess; ++l) ce = cudaMalloc(&n->w[l], wsz[l] * 2);
    if (ce == cudaSuccess) ce = cudaMalloc(&n->w_heads, 32 * 512 * 2);
    if (ce == cudaSuccess) ce = cudaMalloc(&n->w1core, 32 * 32 * 32 * 8 * 2);
    for (int l = 0; l < 7 && ce == cudaSuccess; ++l) ce = cudaMalloc(&n->bias[l], bsz[l] * 4);
    if (ce == cudaSuccess) ce = cudaMalloc(&n->fc1_w, 67 * 5 * 4);
    if (ce == cudaSuccess) ce = cudaMalloc(&n->fc1_b, 67 * 4);
    if (ce == cudaSuccess) ce = cudaMalloc(&n->scalar, 4);
    if (ce == cudaSuccess) ce = cudaMalloc(&n->heads_out, B * 32 * 4);
    for (int l = 0; l < 8 && ce == cudaSuccess; ++l) ce = cudaMalloc(&n->act[l], n->act_elems[l] * 2);
    if (ce == cudaSuccess) ce = cudaMemset(n->w_heads, 0, 32 * 512 * 2);
    if (ce == cudaSuccess) ce = cudaMemset(n->bias[6], 0, 32 * 4);
    if (ce == cudaSuccess) {
        const int smem64 = QG_STAGES * (QG_A_BYTES + 64 * 128) + 256, smem32 = QG_STAGES * (QG_A_BYTES + 32 * 128) + 256;
        ce = cudaFuncSetAttribute(qnet_gemm_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem64);
        if (ce == cudaSuccess) ce = cudaFuncSetAttribute(qnet_gemm_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem32);
        if (ce == cudaSuccess) ce = cudaFuncSetAttribute(qnet_conv1_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, C1_SMEM);
    }
    cudaSetDevice(prev);
    if (ce != cudaSuccess) {
        ffmp_qnet_destroy(n);
        return qfail(FFMP_ERR_CUDA, "qnet allocation", ce);
    }
    *out = n;
    return FFMP_OK;
}

// Weights in the layouts of the reference module's state_dict (train.py:233-242), f32 on the device:
//   conv1 [32][2][32][32], conv2 [64][32][32][32], conv3 / conv4 [64][64][8][8], fc1 [67][5], fc2 [512][6400], fc3 [512][512],
//   fc4_ea [28][512], fc4_ev [1][512]; biases alongside.
int ffmp_qnet_load(ffmp_qnet *n, const float *const *weights, const float *const *biases, void *stream) {
    if (!n || !weights || !biases) return qfail(FFMP_ERR_ARG, "null argument");
    for (int i = 0; i < 9; ++i)
        if (!weights[i] || !biases[i]) return qfail(FFMP_ERR_ARG, "nine weight and nine bias pointers are required");
    int prev = 0;
    cudaGetDevice(&prev);
    cudaSetDevice(n->device);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    // order: conv1, conv2, conv3, conv4, fc1, fc2, fc3, fc4_ea, fc4_ev
    qnet_conv_weight_kernel<<<256, 256, 0, st>>>(weights[0], n->w[0], 32, 2, 32, 32, 8);
    qnet_conv1_weight_core_kernel<<<1024, 256, 0, st>>>(weights[0], n->w1core);
    qnet_conv_weight_kernel<<<1024, 256, 0, st>>>(weights[1], n->w[1], 64, 32, 32, 32, 32);
    qnet_conv_weight_kernel<<<256, 256, 0, st>>>(weights[2], n->w[2], 64, 64, 8, 8, 64);
    qnet_conv_weight_kernel<<<256, 256, 0, st>>>(weights[3], n->w[3], 64, 64, 8, 8, 64);
    qnet_fc2_weight_kernel<<<1024, 256, 0, st>>>(weights[5], n->w[4], 512, 64, 100);
    qnet_cast_rows_kernel<<<256, 256, 0, st>>>(weights[6], n->w[5], 512u * 512);
    qnet_cast_rows_kernel<<<32, 256, 0, st>>>(weights[7], n->w_heads, 28u * 512);
    qnet_cast_rows_kernel<<<2, 256, 0, st>>>(weights[8], n->w_heads + 28 * 512, 512u);
    cudaError_t ce = cudaGetLastError();
    const int bl[4] = {32, 64, 64, 64};
    for (int l = 0; l < 4 && ce == cudaSuccess; ++l) ce = cudaMemcpyAsync(n->bias[l], biases[l], bl[l] * 4, cudaMemcpyDeviceToDevice, st);
    if (ce == cudaSuccess) ce = cudaMemcpyAsync(n->fc1_w, weights[4], 67 * 5 * 4, cudaMemcpyDeviceToDevice, st);
    if (ce == cudaSuccess) ce = cudaMemcpyAsync(n->fc1_b, biases[4], 67 * 4, cudaMemcpyDeviceToDevice, st);
    if (ce == cudaSuccess) ce = cudaMemcpyAsync(n->bias[4], biases[5], 512 * 4, cudaMemcpyDeviceToDevice, st);
    if (ce == cudaSuccess) ce = cudaMemcpyAsync(n->bias[5], biases[6], 512 * 4, cudaMemcpyDeviceToDevice, st);
    if (ce == cudaSuccess) ce = cudaMemcpyAsync(n->bias[6], biases[7], 28 * 4, cudaMemcpyDeviceToDevice, st);
    if (ce == cudaSuccess) ce = cudaMemcpyAsync(n->bias[6] + 28, biases[8], 4, cudaMemcpyDeviceToDevice, st);
    cudaSetDevice(prev);
    if (ce != cudaSuccess) return qfail(FFMP_ERR_CUDA, "ffmp_qnet_load", ce);
    n->loaded = true;
    n->launches += 8;
    return FFMP_OK;
}

// Network.forward(state_m, state_g, state_v, state_t) (train.py:244-303).  state_m: [B][2][100][100] NCHW, bf16 (dtype 1, what
// ffmp_learner_input writes) or f32 (dtype 0); g, v: f32 [B][2]; t: f32 [B][1]; q_out: f32 [B][28].
// scalar_tile != 0 keeps the reference's `x_gvt_[0][30]` scalar tile (train.py:264-276); 0 leaves it out.
int ffmp_qnet_forward(ffmp_qnet *n, int32_t batch, const void *state_m, int32_t dtype, const float *state_g, const float *state_v,
                      const float *state_t, int32_t scalar_tile, float *q_out, void *stream) {
    if (!n || !state_m || !state_g || !state_v || !state_t || !q_out) return qfail(FFMP_ERR_ARG, "null argument");
    if (!n->loaded) return qfail(FFMP_ERR_STATE, "ffmp_qnet_load must be called before ffmp_qnet_forward");
    if (batch <= 0 || batch > n->max_batch) return qfail(FFMP_ERR_ARG, "batch must be in [1, max_batch]");
    if (dtype != 0 && dtype != 1) return qfail(FFMP_ERR_ARG, "dtype must be 0 (float32) or 1 (bfloat16)");
    int prev = 0;
    cudaGetDevice(&prev);
    cudaSetDevice(n->device);
    struct Restore { int d; ~Restore() { cudaSetDevice(d); } } restore{prev};
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int B = batch;
    {
        const size_t px = static_cast<size_t>(B) * 10000;
        const unsigned blocks = static_cast<unsigned>((px + 255) / 256);
        if (dtype == 1) qnet_pack_input_kernel<<<blocks, 256, 0, st>>>(static_cast<const __nv_bfloat16 *>(state_m), n->act[0], B, 10000);
        else qnet_pack_input_kernel<<<blocks, 256, 0, st>>>(static_cast<const float *>(state_m), n->act[0], B, 10000);
        qnet_scalar_tile_kernel<<<1, 32, 0, st>>>(n->fc1_w, n->fc1_b, state_g, state_v, state_t, scalar_tile, n->scalar);
        QCK(cudaGetLastError());
        n->launches += 2;
    }
    // ---- the six convolutions as implicit GEMMs ----
    const bool conv1_resident = !std::getenv("QNET_CONV1_V1");      // development switch: the generic kernel for conv1 (A/B timing)
    if (conv1_resident) {
        qnet_conv1_kernel<<<B, QG_THREADS, C1_SMEM, st>>>(n->act[0], n->w1core, n->bias[0], n->act[1], 0);
        QCK(cudaGetLastError());
        n->launches += 1;
    }
    for (int l = conv1_resident ? 1 : 0; l < 6; ++l) {
        const ConvGeo &c = CONVS[l];
        CUtensorMap ta, tb;
        const cuuint64_t pix = static_cast<cuuint64_t>(c.ICP) * 2;                       // bytes per input pixel
        const cuuint64_t adims[4] = {static_cast<cuuint64_t>(c.KW) * c.ICP, static_cast<cuuint64_t>(c.OW), static_cast<cuuint64_t>(c.IH),
                                     static_cast<cuuint64_t>(B)};
        const cuuint64_t astr[3] = {pix, pix * c.IW, pix * c.IW * c.IH};
        const cuuint32_t abox[4] = {64, static_cast<cuuint32_t>(c.OW), static_cast<cuuint32_t>(c.rows), 1};
        const int K = c.KH * c.KW * c.ICP;
        const cuuint64_t bdims[2] = {static_cast<cuuint64_t>(K), static_cast<cuuint64_t>(c.OC)};
        const cuuint64_t bstr[1] = {static_cast<cuuint64_t>(K) * 2};
        const cuuint32_t bbox[2] = {64, static_cast<cuuint32_t>(c.OC)};
        const int wi = l < 3 ? l : 3;                                                    // conv4's weights serve three layers
        if (!make_map(&ta, n->act[l], 4, adims, astr, abox) || !make_map(&tb, n->w[wi], 2, bdims, bstr, bbox))
            return qfail(FFMP_ERR_CUDA, "cuTensorMapEncodeTiled failed for a convolution");
        QGemmArgs g{};
        g.k_outer = c.KH; g.k_inner = c.KW * c.ICP / 64;
        g.a_c1_step = 0; g.a_c2_step = c.rows;
        g.box1 = c.OW; g.box2 = c.rows; g.lim1 = c.OW; g.lim2 = c.OH;
        g.ldc = c.OC; g.relu = 1; g.out_f32 = 0;
        g.bias = n->bias[wi];
        g.post_add = l == 2 ? n->scalar : nullptr;                                       // x_pls = relu(conv3) + tile
        g.out = n->act[l + 1];
        const dim3 grid(static_cast<unsigned>((c.OH + c.rows - 1) / c.rows), static_cast<unsigned>(B), 1);
        if (c.OC == 32) qnet_gemm_kernel<32><<<grid, QG_THREADS, QG_STAGES * (QG_A_BYTES + 32 * 128) + 256, st>>>(ta, tb, g);
        else qnet_gemm_kernel<64><<<grid, QG_THREADS, QG_STAGES * (QG_A_BYTES + 64 * 128) + 256, st>>>(ta, tb, g);
        QCK(cudaGetLastError());
        n->launches += 1;
    }
    // ---- fc2, fc3, heads: the same kernel on a rank-2 activation ----
    struct Fc { const __nv_bfloat16 *in; const __nv_bfloat16 *w; const float *bias; void *out; int K, N, relu, f32; };
    __nv_bfloat16 *fc2_out = n->act[7], *fc3_out = n->act[7] + static_cast<size_t>(n->max_batch) * 512;
    const Fc fcs[3] = {{n->act[6], n->w[4], n->bias[4], fc2_out, 6400, 512, 1, 0},
                       {fc2_out, n->w[5], n->bias[5], fc3_out, 512, 512, 1, 0},
                       {fc3_out, n->w_heads, n->bias[6], n->heads_out, 512, 32, 0, 1}};
    for (const Fc &f : fcs) {
        CUtensorMap ta, tb;
        const cuuint64_t adims[4] = {static_cast<cuuint64_t>(f.K), static_cast<cuuint64_t>(B), 1, 1};
        const cuuint64_t astr[3] = {static_cast<cuuint64_t>(f.K) * 2, static_cast<cuuint64_t>(f.K) * 2 * B, static_cast<cuuint64_t>(f.K) * 2 * B};
        const cuuint32_t abox[4] = {64, 128, 1, 1};
        const cuuint64_t bdims[2] = {static_cast<cuuint64_t>(f.K), static_cast<cuuint64_t>(f.N)};
        const cuuint64_t bstr[1] = {static_cast<cuuint64_t>(f.K) * 2};
        const cuuint32_t bbox[2] = {64, static_cast<cuuint32_t>(f.N < 64 ? f.N : 64)};
        if (!make_map(&ta, const_cast<__nv_bfloat16 *>(f.in), 4, adims, astr, abox) ||
            !make_map(&tb, const_cast<__nv_bfloat16 *>(f.w), 2, bdims, bstr, bbox))
            return qfail(FFMP_ERR_CUDA, "cuTensorMapEncodeTiled failed for a fully connected layer");
        QGemmArgs g{};
        g.k_outer = 1; g.k_inner = f.K / 64;
        g.a_c1_step = 128; g.a_c2_step = 0;
        g.box1 = 128; g.box2 = 1; g.lim1 = B; g.lim2 = 1;
        g.ldc = f.N; g.relu = f.relu; g.out_f32 = f.f32;
        g.bias = f.bias; g.out = f.out;
        if (const char *e = std::getenv("QNET_DBG_SHIFT")) g.dbg_shift = std::atoi(e);
        if (const char *e = std::getenv("QNET_DBG_BO")) g.dbg_bo = std::atoi(e);
        const dim3 grid(static_cast<unsigned>((B + 127) / 128), 1, static_cast<unsigned>(f.N < 64 ? 1 : f.N / 64));
        if (f.N < 64) qnet_gemm_kernel<32><<<grid, QG_THREADS, QG_STAGES * (QG_A_BYTES + 32 * 128) + 256, st>>>(ta, tb, g);
        else qnet_gemm_kernel<64><<<grid, QG_THREADS, QG_STAGES * (QG_A_BYTES + 64 * 128) + 256, st>>>(ta, tb, g);
        QCK(cudaGetLastError());
        n->launches += 1;
    }
    qnet_dueling_kernel<<<static_cast<unsigned>((B + 3) / 4), 128, 0, st>>>(n->heads_out, q_out, B);
    QCK(cudaGetLastError());
    n->launches += 1;
    return FFMP_OK;
}

// Intermediate activations for layer-by-layer tests: layer 1..6 = the six convolution outputs (bf16 NHWC), 7 = fc2, 8 = fc3
// (bf16 [B][512]); copies batch * elems bf16 values to out_dev.
int ffmp_qnet_debug_activation(ffmp_qnet *n, int32_t layer, int32_t batch, void *out_dev, size_t *elems_per_sample, void *stream) {
    if (!n || layer < 1 || layer > 8 || batch <= 0 || batch > n->max_batch) return qfail(FFMP_ERR_ARG, "bad argument");
    size_t per = 512;
    const __nv_bfloat16 *src;
    if (layer <= 6) { per = static_cast<size_t>(CONVS[layer - 1].OH) * CONVS[layer - 1].OW * CONVS[layer - 1].OC; src = n->act[layer]; }
    else src = n->act[7] + (layer == 8 ? static_cast<size_t>(n->max_batch) * 512 : 0);
    if (elems_per_sample) *elems_per_sample = per;
    if (out_dev) QCK(cudaMemcpyAsync(out_dev, src, per * batch * 2, cudaMemcpyDeviceToDevice, static_cast<cudaStream_t>(stream)));
    return FFMP_OK;
}

int ffmp_qnet_launch_count(const ffmp_qnet *n, uint64_t *out) {
    if (!n || !out) return qfail(FFMP_ERR_ARG, "null argument");
    *out = n->launches;
    return FFMP_OK;
}

}  // extern "C"
