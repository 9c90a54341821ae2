// host_io.cu — the host-buffer side of ffmp_step_host: the per-step results go straight into the caller's pinned host
// memory from a kernel, and the host learns about completion from a flag word in mapped memory instead of a stream
// synchronisation.  A DMA copy + cudaStreamSynchronize costs two engine hand-overs and a driver wake-up per step
// (~20 us, as much as the env step itself); this path costs one small dependent kernel.
//
// host_export_kernel is launched right behind the tick kernel with programmatic stream serialization: its single CTA is
// resident while the tick's last wave runs, `griddepcontrol.wait` returns when the tick grid has completed and its
// writes are visible, then 1024 threads move the packed result block (reward | rel_goal | velocity | done | flags,
// 22 N bytes) with 16-byte accesses over PCIe, fence at system scope and publish the step's sequence number.
#include "ffmp_kernels.cuh"

namespace ffmp {

namespace {

__global__ void __launch_bounds__(1024) host_export_kernel(HostExportArgs a) {
    unsigned long long g0 = 0, g1 = 0;
    if (a.stamps && threadIdx.x == 0) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g0));
    asm volatile("griddepcontrol.wait;" ::: "memory");
    if (a.stamps && threadIdx.x == 0) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g1));
    const uint4 *src = reinterpret_cast<const uint4 *>(a.src);
    uint4 *dst = reinterpret_cast<uint4 *>(a.dst);
    constexpr int UN = 6;   // 22 * 4096 B = 5632 x 16 B: every load of the common batch size is in flight at once
    for (int k0 = threadIdx.x; k0 < a.n16; k0 += UN * 1024) {
        uint4 x[UN];
#pragma unroll
        for (int i = 0; i < UN; ++i)
            if (k0 + i * 1024 < a.n16) x[i] = __ldcg(src + k0 + i * 1024);
#pragma unroll
        for (int i = 0; i < UN; ++i)
            if (k0 + i * 1024 < a.n16) dst[k0 + i * 1024] = x[i];
    }
    if (static_cast<int>(threadIdx.x) < a.tail_words) {
        const uint32_t *s = reinterpret_cast<const uint32_t *>(src + a.n16);
        uint32_t *d = reinterpret_cast<uint32_t *>(dst + a.n16);
        d[threadIdx.x] = __ldcg(s + threadIdx.x);
    }
    // The CTA barrier orders every thread's stores before thread 0's system-scope fence (cumulativity), so ONE fence
    // publishes the whole block; 32 warps fencing at system scope one after the other cost microseconds.
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence_system();
        if (a.stamps) {   // diagnostics (FFMP_HOST_IO_STATS=1): resident / grid dependency released / block written
            unsigned long long g2;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g2));
            a.stamps[0] = g0; a.stamps[1] = g1; a.stamps[2] = g2;
            __threadfence_system();
        }
        *reinterpret_cast<volatile uint32_t *>(a.flag) = a.value;
    }
}

}  // namespace

cudaError_t launch_host_export(const HostExportArgs &a, cudaStream_t st) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(1);
    cfg.blockDim = dim3(1024);
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, host_export_kernel, a);
}

}  // namespace ffmp
