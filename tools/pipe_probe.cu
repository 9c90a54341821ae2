// pipe_probe.cu — issue-rate probe for the integer instruction mixes of the bit-parallel wavefront
// (which pipe LOP3 / SHF / IMAD / IMAD.WIDE / IMAD.HI / PRMT / SHFL go to, and how well they overlap).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/pipe_probe tools/pipe_probe.cu && /tmp/pipe_probe
// Prints warp-instructions per clock per SM for each mix (16 warps per SMSP resident, 8 independent chains).
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("ERR %s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

constexpr int CH = 8;       // independent chains per thread
constexpr int ITER = 4096;

template <int MODE>
__global__ void __launch_bounds__(256) probe(uint32_t *out, uint32_t m0, uint32_t m1, uint32_t m2) {
    uint32_t a[CH], b[CH];
#pragma unroll
    for (int c = 0; c < CH; ++c) { a[c] = threadIdx.x * 2654435761u + c; b[c] = blockIdx.x + c * 40503u; }
#pragma unroll 1
    for (int it = 0; it < ITER; ++it) {
#pragma unroll
        for (int c = 0; c < CH; ++c) {
            if (MODE == 0) {            // LOP3 only (2 per chain)
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[c]) : "r"(b[c]), "r"(m0));
                asm volatile("lop3.b32 %0, %0, %1, %2, 0xe8;" : "+r"(b[c]) : "r"(a[c]), "r"(m1));
            } else if (MODE == 1) {     // SHF only
                asm volatile("shf.l.wrap.b32 %0, %0, %1, %2;" : "+r"(a[c]) : "r"(b[c]), "r"(m2));
                asm volatile("shf.r.wrap.b32 %0, %0, %1, %2;" : "+r"(b[c]) : "r"(a[c]), "r"(m2));
            } else if (MODE == 2) {     // IMAD only
                asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a[c]) : "r"(m0), "r"(b[c]));
                asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(b[c]) : "r"(m1), "r"(a[c]));
            } else if (MODE == 3) {     // IMAD.WIDE only
                uint64_t t, u;
                asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(t) : "r"(a[c]), "r"(m0));
                asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(u) : "r"(b[c]), "r"(m1));
                a[c] = static_cast<uint32_t>(t) ^ 0u; b[c] = static_cast<uint32_t>(u >> 32);
                a[c] += static_cast<uint32_t>(t >> 32); b[c] += static_cast<uint32_t>(u);
            } else if (MODE == 4) {     // IMAD.HI only
                asm volatile("mad.hi.u32 %0, %0, %1, %2;" : "+r"(a[c]) : "r"(m0), "r"(b[c]));
                asm volatile("mad.hi.u32 %0, %0, %1, %2;" : "+r"(b[c]) : "r"(m1), "r"(a[c]));
            } else if (MODE == 5) {     // LOP3 + IMAD 1:1
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[c]) : "r"(b[c]), "r"(m0));
                asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(b[c]) : "r"(m1), "r"(a[c]));
            } else if (MODE == 6) {     // LOP3 + SHF 1:1 (same pipe?)
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[c]) : "r"(b[c]), "r"(m0));
                asm volatile("shf.l.wrap.b32 %0, %0, %1, %2;" : "+r"(b[c]) : "r"(a[c]), "r"(m2));
            } else if (MODE == 7) {     // PRMT only
                asm volatile("prmt.b32 %0, %0, %1, %2;" : "+r"(a[c]) : "r"(b[c]), "r"(m0));
                asm volatile("prmt.b32 %0, %0, %1, %2;" : "+r"(b[c]) : "r"(a[c]), "r"(m1));
            } else if (MODE == 8) {     // 2 LOP3 + 1 IMAD
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[c]) : "r"(b[c]), "r"(m0));
                asm volatile("lop3.b32 %0, %0, %1, %2, 0xe8;" : "+r"(b[c]) : "r"(a[c]), "r"(m1));
                asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a[c]) : "r"(m1), "r"(b[c]));
            } else if (MODE == 9) {     // LOP3 + FADD 1:1 (fp32 pipe beside the integer one)
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[c]) : "r"(b[c]), "r"(m0));
                float f = __uint_as_float(b[c]);
                asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(f) : "f"(__uint_as_float(m1)));
                b[c] = __float_as_uint(f);
            } else if (MODE == 10) {    // LOP3 + IMAD + FADD 1:1:1
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[c]) : "r"(b[c]), "r"(m0));
                asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(b[c]) : "r"(m1), "r"(a[c]));
                float f = __uint_as_float(b[c]);
                asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(f) : "f"(__uint_as_float(m1)));
                b[c] = __float_as_uint(f);
            } else if (MODE == 11) {    // SHFL only
                a[c] = __shfl_up_sync(0xFFFFFFFFu, a[c], 1);
                b[c] = __shfl_down_sync(0xFFFFFFFFu, b[c], 1);
            } else if (MODE == 12) {    // IMAD.WIDE + LOP3 1:1
                uint64_t t;
                asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(t) : "r"(a[c]), "r"(m0));
                asm volatile("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(a[c]) : "r"(static_cast<uint32_t>(t)), "r"(static_cast<uint32_t>(t >> 32)), "r"(b[c]));
            } else if (MODE == 13) {    // IADD3 only
                asm volatile("add.u32 %0, %0, %1;" : "+r"(a[c]) : "r"(b[c]));
                asm volatile("add.u32 %0, %0, %1;" : "+r"(b[c]) : "r"(a[c]));
            } else if (MODE == 14) {    // SEL only (selp)
                asm volatile("{.reg .pred p; setp.ne.u32 p, %1, 0; selp.u32 %0, %0, %2, p;}" : "+r"(a[c]) : "r"(b[c]), "r"(m0));
                asm volatile("{.reg .pred p; setp.ne.u32 p, %1, 0; selp.u32 %0, %0, %2, p;}" : "+r"(b[c]) : "r"(a[c]), "r"(m1));
            }
        }
    }
    uint32_t s = 0;
#pragma unroll
    for (int c = 0; c < CH; ++c) s ^= a[c] ^ b[c];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
int run(const char *name, double inst_per_iter_chain, uint32_t *out, int sms, double clk_ghz) {
    const int grid = sms * 8;   // 8 CTAs x 8 warps = 64 warps per SM
    probe<MODE><<<grid, 256>>>(out, 0x9E3779B9u, 0x85EBCA6Bu, 7u);
    CK(cudaDeviceSynchronize());
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    CK(cudaEventRecord(e0));
    probe<MODE><<<grid, 256>>>(out, 0x9E3779B9u, 0x85EBCA6Bu, 7u);
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    const double winst = static_cast<double>(grid) * 8 * ITER * CH * inst_per_iter_chain;
    const double per_clk_sm = winst / (ms * 1e-3 * clk_ghz * 1e9) / sms;
    printf("%-28s %8.3f ms  %6.2f warp-inst/clk/SM (counted %.0f inst/iter/chain)\n", name, ms, per_clk_sm, inst_per_iter_chain);
    return 0;
}

int main() {
    cudaDeviceProp p;
    CK(cudaGetDeviceProperties(&p, 0));
    int khz = 0;
    CK(cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0));
    const double ghz = khz * 1e-6;
    printf("%s, %d SMs, %.3f GHz (max clock; rates assume it)\n", p.name, p.multiProcessorCount, ghz);
    uint32_t *out;
    CK(cudaMalloc(&out, static_cast<size_t>(p.multiProcessorCount) * 8 * 256 * 4));
    const int sms = p.multiProcessorCount;
    run<0>("LOP3", 2, out, sms, ghz);
    run<1>("SHF", 2, out, sms, ghz);
    run<2>("IMAD", 2, out, sms, ghz);
    run<3>("IMAD.WIDE (+2 IADD)", 4, out, sms, ghz);
    run<4>("IMAD.HI", 2, out, sms, ghz);
    run<5>("LOP3+IMAD 1:1", 2, out, sms, ghz);
    run<6>("LOP3+SHF 1:1", 2, out, sms, ghz);
    run<7>("PRMT", 2, out, sms, ghz);
    run<8>("2 LOP3 + 1 IMAD", 3, out, sms, ghz);
    run<9>("LOP3+FADD 1:1", 2, out, sms, ghz);
    run<10>("LOP3+IMAD+FADD", 3, out, sms, ghz);
    run<11>("SHFL", 2, out, sms, ghz);
    run<12>("IMAD.WIDE+LOP3 1:1", 2, out, sms, ghz);
    run<13>("IADD", 2, out, sms, ghz);
    run<14>("ISETP+SEL", 4, out, sms, ghz);
    return 0;
}
