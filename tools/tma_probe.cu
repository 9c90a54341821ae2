// parametrised 2-D probe: elem size 1 or 4, global width, box
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda/barrier>
#include <cstdio>
#include <cstdlib>
#include <vector>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("ERR %s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)
using barrier = cuda::barrier<cuda::thread_scope_block>;
namespace cde = cuda::device::experimental;
__global__ void probe(const __grid_constant__ CUtensorMap tmap, int c0, int c1, int bytes, uint8_t *out) {
    extern __shared__ __align__(128) uint8_t tile[];
#pragma nv_diag_suppress static_var_with_dynamic_init
    __shared__ barrier bar;
    if (threadIdx.x == 0) { init(&bar, blockDim.x); cde::fence_proxy_async_shared_cta(); }
    __syncthreads();
    barrier::arrival_token token;
    if (threadIdx.x == 0) {
        cde::cp_async_bulk_tensor_2d_global_to_shared(tile, &tmap, c0, c1, bar);
        token = cuda::device::barrier_arrive_tx(bar, 1, bytes);
    } else token = bar.arrive();
    bar.wait(std::move(token));
    for (int k = threadIdx.x; k < bytes; k += blockDim.x) out[k] = tile[k];
}
int main(int argc, char **argv) {
    const int es = atoi(argv[1]), GW = atoi(argv[2]), GH = atoi(argv[3]), BW = atoi(argv[4]), BH = atoi(argv[5]);
    const int c0 = argc > 6 ? atoi(argv[6]) : 0, c1 = argc > 7 ? atoi(argv[7]) : 0;
    std::vector<uint8_t> h((size_t)GW * GH * es);
    for (size_t i = 0; i < h.size(); ++i) h[i] = (uint8_t)(i * 2654435761u >> 24);
    uint8_t *d, *o;
    const int bytes = BW * BH * es;
    CK(cudaMalloc(&d, h.size())); CK(cudaMalloc(&o, bytes));
    CK(cudaMemcpy(d, h.data(), h.size(), cudaMemcpyHostToDevice));
    typedef CUresult (*EncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                 const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                 CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    void *fn = nullptr; cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
    CUtensorMap tm{};
    cuuint64_t dims[2] = {(cuuint64_t)GW, (cuuint64_t)GH}; cuuint64_t strides[1] = {(cuuint64_t)GW * es};
    cuuint32_t box[2] = {(cuuint32_t)BW, (cuuint32_t)BH}; cuuint32_t est[2] = {1, 1};
    CUresult r = ((EncodeFn)fn)(&tm, es == 1 ? CU_TENSOR_MAP_DATA_TYPE_UINT8 : CU_TENSOR_MAP_DATA_TYPE_INT32, 2, d, dims, strides, box, est,
                                CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("es=%d G=%dx%d box=%dx%d at (%d,%d): encode r=%d ", es, GW, GH, BW, BH, c0, c1, (int)r);
    if (r) { printf("\n"); return 2; }
    CK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    probe<<<1, 128, bytes>>>(tm, c0, c1, bytes, o);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("RUN ERR: %s\n", cudaGetErrorString(e)); return 1; }
    std::vector<uint8_t> got(bytes);
    CK(cudaMemcpy(got.data(), o, got.size(), cudaMemcpyDeviceToHost));
    int bad = 0;
    for (int a = 0; a < BH; ++a) for (int b = 0; b < BW * es; ++b) {
        int i = c1 + a, j = c0 * es + b;
        uint8_t ex = (i < 0 || j < 0 || i >= GH || j >= GW * es) ? 0 : h[(size_t)i * GW * es + j];
        bad += got[a * BW * es + b] != ex;
    }
    printf("-> %d mismatches\n", bad);
    return 0;
}
