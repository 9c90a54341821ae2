// tmap_probe.cu — does cuTensorMapEncodeTiled accept OVERLAPPING strides (a sliding window expressed as a dimension), and does
// the copy deliver what the strides say?  (qnet.cu's implicit-GEMM convolutions rely on it for the 32 x 32 kernels.)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tools/tmap_probe tools/tmap_probe.cu && tools/tmap_probe
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <vector>

typedef CUresult (*EncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                             const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

__global__ void probe(const __grid_constant__ CUtensorMap tm, int c0, int c1, int c2, int c3, uint16_t *out, int bytes) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar;
    const uint32_t b = static_cast<uint32_t>(__cvta_generic_to_shared(&bar));
    const uint32_t d = static_cast<uint32_t>(__cvta_generic_to_shared(smem));
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
                     ::"r"(d), "l"(&tm), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(b) : "memory");
    }
    __syncthreads();
    uint32_t done = 0;
    while (!done)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(b), "r"(0u) : "memory");
    for (int i = threadIdx.x; i < bytes / 2; i += blockDim.x) out[i] = reinterpret_cast<uint16_t *>(smem)[i];
}

int main() {
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaFree(0);
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess || !fn) { printf("no entry point\n"); return 1; }
    // conv2 of the Q network: input NHWC [N=2][H=69][W=69][C=32] u16 (bf16 bit patterns = the flat element index mod 65536)
    const int N = 2, H = 69, W = 69, C = 32, OW = 38;
    std::vector<uint16_t> h(static_cast<size_t>(N) * H * W * C);
    for (size_t i = 0; i < h.size(); ++i) h[i] = static_cast<uint16_t>(i % 65521);
    uint16_t *dptr, *dout;
    cudaMalloc(&dptr, h.size() * 2);
    cudaMemcpy(dptr, h.data(), h.size() * 2, cudaMemcpyHostToDevice);
    cudaMalloc(&dout, 32768);
    // dims: d0 = (kw, c) window inside a row [32 * 32], d1 = ow [38] with stride C elements (OVERLAPS d0), d2 = input row, d3 = sample
    const cuuint64_t dims[4] = {1024, OW, H, N};
    const cuuint64_t strides[3] = {C * 2, static_cast<cuuint64_t>(W) * C * 2, static_cast<cuuint64_t>(H) * W * C * 2};
    const cuuint32_t box[4] = {64, OW, 3, 1};
    const cuuint32_t es[4] = {1, 1, 1, 1};
    CUtensorMap tm;
    CUresult r = reinterpret_cast<EncodeFn>(fn)(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, dptr, dims, strides, box, es,
                                                CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                                CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode (overlapping strides, 128B swizzle) -> %d\n", static_cast<int>(r));
    if (r != CUDA_SUCCESS) return 2;
    const int bytes = 64 * OW * 3 * 2;
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768);
    const int kw0 = 6, ih0 = 5, n = 1;
    probe<<<1, 128, 32768>>>(tm, kw0 * C, 0, ih0, n, dout, bytes);
    cudaError_t ce = cudaDeviceSynchronize();
    printf("kernel -> %s\n", cudaGetErrorString(ce));
    if (ce != cudaSuccess) return 3;
    std::vector<uint16_t> o(bytes / 2);
    cudaMemcpy(o.data(), dout, bytes, cudaMemcpyDeviceToHost);
    // expected: row p = (r * 38 + ow), 64 elements = x[n][ih0 + r][ow + kw0 .. +1][0..31]; 128B swizzle: 16-byte chunk j of row p at j ^ (p & 7)
    int bad = 0;
    for (int rr = 0; rr < 3; ++rr)
        for (int ow = 0; ow < OW; ++ow)
            for (int k = 0; k < 64; ++k) {
                const int p = rr * OW + ow;
                const size_t src = ((static_cast<size_t>(n) * H + ih0 + rr) * W + ow + kw0) * C + k;
                const int chunk = (k / 8) ^ (p & 7);
                const uint16_t got = o[p * 64 + chunk * 8 + (k & 7)];
                if (got != static_cast<uint16_t>(src % 65521)) ++bad;
            }
    printf("mismatches: %d of %d\n", bad, 3 * OW * 64);
    return bad ? 4 : 0;
}
