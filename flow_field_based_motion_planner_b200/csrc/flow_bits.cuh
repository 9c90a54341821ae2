// flow_bits.cuh — bit-plane helpers shared by the flow-field kernels (flow_field.cu, flow_field_large.cu): vector row
// accesses, funnel shifts across row words, PRMT / bit-matrix transposes, TMA bulk copy + mbarrier wrappers, and the
// FMA-pipe forms of the few non-logic operations of the wavefront loop.
#pragma once
#include "ffmp_kernels.cuh"

namespace ffmp {
namespace {

template <int WPR>
struct Row {
    static __device__ __forceinline__ void ld(const uint32_t *p, uint32_t (&v)[WPR]) {
        if constexpr (WPR % 4 == 0) {
#pragma unroll
            for (int q = 0; q < WPR / 4; ++q) {
                const uint4 t = reinterpret_cast<const uint4 *>(p)[q];
                v[4 * q] = t.x; v[4 * q + 1] = t.y; v[4 * q + 2] = t.z; v[4 * q + 3] = t.w;
            }
        } else if constexpr (WPR == 2) {
            const uint2 t = *reinterpret_cast<const uint2 *>(p);
            v[0] = t.x; v[1] = t.y;
        } else {
#pragma unroll
            for (int w = 0; w < WPR; ++w) v[w] = p[w];
        }
    }
    static __device__ __forceinline__ void st(uint32_t *p, const uint32_t (&v)[WPR]) {
        if constexpr (WPR % 4 == 0) {
#pragma unroll
            for (int q = 0; q < WPR / 4; ++q)
                reinterpret_cast<uint4 *>(p)[q] = make_uint4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
        } else if constexpr (WPR == 2) {
            *reinterpret_cast<uint2 *>(p) = make_uint2(v[0], v[1]);
        } else {
#pragma unroll
            for (int w = 0; w < WPR; ++w) p[w] = v[w];
        }
    }
};

// spread the 4 bits of nibble n of x to the low bit of 4 bytes
__device__ __forceinline__ uint32_t spread4(uint32_t x, int n) {
    return (((x >> (4 * n)) & 0xFu) * 0x00204081u) & 0x01010101u;
}

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// End of a flow-field launch, called by ONE thread per CTA after the CTA's last grid (the CTA barrier of the grid loop orders the
// other threads' stores before it): fence, take a ticket; the LAST CTA re-arms the ticket, the dynamic hand-out counter and the
// regeneration list's count for their next use and, for background regenerations, publishes the launch's sequence number in
// mapped host memory (every CTA fenced its writes before it took its ticket, so the launch's results are visible by then).
__device__ __forceinline__ void flow_launch_epilogue(const FlowArgs &a) {
    __threadfence();
    const uint32_t t = atomicAdd(a.ticket, 1u);
    if (t == gridDim.x - 1) {
        *a.ticket = 0;
        if (a.work) *a.work = 0;
        if (a.count_reset) *a.count_reset = 0;
        __threadfence();
        if (a.host_done) {
            __threadfence_system();
            *reinterpret_cast<volatile uint32_t *>(a.host_done) = a.host_done_value;
        }
    }
}

// 8x8 bit-matrix transpose of the 64-bit value hi:lo (Hacker's Delight 7-3), on 32-bit halves
__device__ __forceinline__ void transpose8(uint32_t &lo, uint32_t &hi) {
    uint32_t t;
    t = (lo ^ (lo >> 7)) & 0x00AA00AAu; lo ^= t ^ (t << 7);
    t = (hi ^ (hi >> 7)) & 0x00AA00AAu; hi ^= t ^ (t << 7);
    t = (lo ^ (lo >> 14)) & 0x0000CCCCu; lo ^= t ^ (t << 14);
    t = (hi ^ (hi >> 14)) & 0x0000CCCCu; hi ^= t ^ (t << 14);
    t = (lo ^ __funnelshift_r(lo, hi, 28)) & 0xF0F0F0F0u;
    lo ^= t; hi ^= t >> 4;
}

// 4x4 byte transpose: out[b] = (p0.b, p1.b, p2.b, p3.b)
__device__ __forceinline__ void bytes4x4(uint32_t p0, uint32_t p1, uint32_t p2, uint32_t p3, uint32_t (&out)[4]) {
    const uint32_t a = __byte_perm(p0, p1, 0x5140), b = __byte_perm(p0, p1, 0x7362);
    const uint32_t c = __byte_perm(p2, p3, 0x5140), d = __byte_perm(p2, p3, 0x7362);
    out[0] = __byte_perm(a, c, 0x5410); out[1] = __byte_perm(a, c, 0x7632);
    out[2] = __byte_perm(b, d, 0x5410); out[3] = __byte_perm(b, d, 0x7632);
}

// Neighbour columns of a row of WPR words: bit j of the result <- the cell at column-1 (from_lo) / column+1 (from_hi)
// of the cell that bit j of word w stands for.
//   linear layout (IL = false): word w holds columns 32w..32w+31 -> funnel shifts across the row words;
//   column-interleaved layout (IL = true, WPR = 4): word w holds the columns c with c % 4 == w, bit b <-> column 4b + w.
//   The horizontal neighbours of word w are then simply words w-1 / w+1 (register renames, no instruction); only the
//   wrap-around words need a plain shift: 2 shifts per row instead of 8 funnel shifts.
template <int WPR, bool IL = false>
__device__ __forceinline__ uint32_t from_lo(const uint32_t (&x)[WPR], int w) {   // bit j <- cell j-1
    if constexpr (IL) return w > 0 ? x[w - 1] : x[WPR - 1] << 1;
    else return w > 0 ? __funnelshift_l(x[w - 1], x[w], 1) : x[w] << 1;
}
template <int WPR, bool IL = false>
__device__ __forceinline__ uint32_t from_hi(const uint32_t (&x)[WPR], int w) {   // bit j <- cell j+1
    if constexpr (IL) return w + 1 < WPR ? x[w + 1] : x[0] >> 1;
    else return w + 1 < WPR ? __funnelshift_r(x[w], x[w + 1], 1) : x[w] >> 1;
}

// one delta swap: exchange the bits selected by m with the bits d positions above them
__device__ __forceinline__ uint32_t delta_swap(uint32_t x, uint32_t m, int d) {
    const uint32_t t = (x ^ (x >> d)) & m;
    return x ^ t ^ (t << d);
}
// 8 nibbles -> 4 bytes: bit 4q + w -> bit 8w + q (index-bit rotation as four delta swaps), and its inverse
__device__ __forceinline__ uint32_t nibbles_to_bytes(uint32_t x) {
    x = delta_swap(x, 0x22222222u, 1); x = delta_swap(x, 0x0A0A0A0Au, 3);
    x = delta_swap(x, 0x00CC00CCu, 6); return delta_swap(x, 0x0000F0F0u, 12);
}
__device__ __forceinline__ uint32_t bytes_to_nibbles(uint32_t x) {
    x = delta_swap(x, 0x0000F0F0u, 12); x = delta_swap(x, 0x00CC00CCu, 6);
    x = delta_swap(x, 0x0A0A0A0Au, 3); return delta_swap(x, 0x22222222u, 1);
}
// a - n (n subset of a) and x * m (m in {0,1}) forced onto the FMA pipe (IMAD) so that they do not compete with
// the LOP3/SHF stream on the ALU pipe: the multipliers are runtime values ptxas cannot fold into IADD/LOP3/SEL.
__device__ __forceinline__ uint32_t sub_on_fma(uint32_t a, uint32_t n, uint32_t neg1) {
    uint32_t r;
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(n), "r"(neg1), "r"(a));
    return r;
}
__device__ __forceinline__ uint32_t mask_on_fma(uint32_t x, uint32_t m01) {
    uint32_t r;
    asm("mul.lo.u32 %0, %1, %2;" : "=r"(r) : "r"(x), "r"(m01));
    return r;
}
// replicate the sign bit of every byte over that byte (0x80.. -> 0xFF, else 0x00): PRMT with the msb of each selector nibble
__device__ __forceinline__ uint32_t byte_sign_fill(uint32_t x) {
    uint32_t r;
    asm("prmt.b32 %0, %1, %1, 0xba98;" : "=r"(r) : "r"(x));
    return r;
}

// 32 cells x (d0..d3 direction-code planes, occupied) -> 32 flow bytes (255 occupied, else code*28), 8 words
__device__ __forceinline__ void flow_bytes32(uint32_t d0, uint32_t d1, uint32_t d2, uint32_t d3, uint32_t occ, uint32_t (&out)[8]) {
    uint32_t tl[4];
    bytes4x4(d0, d1, d2, d3, tl);
#pragma unroll
    for (int b = 0; b < 4; ++b) {
        uint32_t lo = tl[b], hb = ((occ >> (8 * b)) & 0xFFu) << 24;
        transpose8(lo, hb);          // byte j of hb:lo = code of cell 8b+j, bit 7 = occupied
        out[2 * b] = ((lo & 0x0F0F0F0Fu) * 28u) | byte_sign_fill(lo);
        out[2 * b + 1] = ((hb & 0x0F0F0F0Fu) * 28u) | byte_sign_fill(hb);
    }
}

// A row of 128 cells: linear words (word j = columns 32j..32j+31) <-> column-interleaved words (word w bit b = column 4b + w).
// Interleaved word w, byte j = the bits w, w+4, .., w+28 of linear word j.
__device__ __forceinline__ void row_to_interleaved(uint32_t (&x)[4]) {
    uint32_t o[4];
    bytes4x4(nibbles_to_bytes(x[0]), nibbles_to_bytes(x[1]), nibbles_to_bytes(x[2]), nibbles_to_bytes(x[3]), o);
#pragma unroll
    for (int w = 0; w < 4; ++w) x[w] = o[w];
}
__device__ __forceinline__ void row_to_linear(uint32_t (&x)[4]) {
    uint32_t o[4];
    bytes4x4(x[0], x[1], x[2], x[3], o);        // the byte transpose is its own inverse
#pragma unroll
    for (int j = 0; j < 4; ++j) x[j] = bytes_to_nibbles(o[j]);
}

// Interleaved row (4 words per plane) x (d0..d3 direction-code planes, occupied) -> the 128 flow bytes of the row in column
// order, 32 words: byte planes -> per-word cell bytes (PRMT + 8x8 bit transposes) -> 4x4 byte transposes across the words.
__device__ __forceinline__ void flow_bytes128_il(const uint32_t (&d0)[4], const uint32_t (&d1)[4], const uint32_t (&d2)[4],
                                                 const uint32_t (&d3)[4], const uint32_t (&occ)[4], uint32_t (&out)[32]) {
    uint32_t tl[4][4];
#pragma unroll
    for (int w = 0; w < 4; ++w) bytes4x4(d0[w], d1[w], d2[w], d3[w], tl[w]);
#pragma unroll
    for (int b = 0; b < 4; ++b) {
        uint32_t lo[4], hb[4];
#pragma unroll
        for (int w = 0; w < 4; ++w) {
            lo[w] = tl[w][b]; hb[w] = ((occ[w] >> (8 * b)) & 0xFFu) << 24;
            transpose8(lo[w], hb[w]);      // byte j of hb:lo = code of the cell at bit 8b+j of word w, bit 7 = occupied
        }
        uint32_t x[4], y[4];
        bytes4x4(lo[0], lo[1], lo[2], lo[3], x);   // x[j] = columns 4(8b+j) .. +3
        bytes4x4(hb[0], hb[1], hb[2], hb[3], y);   // y[j] = columns 4(8b+4+j) .. +3
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            out[8 * b + j] = ((x[j] & 0x0F0F0F0Fu) * 28u) | byte_sign_fill(x[j]);
            out[8 * b + 4 + j] = ((y[j] & 0x0F0F0F0Fu) * 28u) | byte_sign_fill(y[j]);
        }
    }
}

}  // namespace
}  // namespace ffmp
