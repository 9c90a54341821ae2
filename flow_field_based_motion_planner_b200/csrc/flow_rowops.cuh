// flow_rowops.cuh — per-row bit-plane arithmetic of the interleaved flow-field kernel (flow_field_il.cu), written so that the
// same source compiles for the device (PRMT / LOP3 / IMAD) and for the host (tests/host/rowops_host.cpp runs every function
// against a per-cell restatement without a GPU).  SPEC.md §4 (F1 cost, F2 direction) and §5 (flow image) are what the
// functions implement; the reference has no flow-field code (the consumer of the image is /root/reference/src/train.py:116-121).
//
// Row layout ("interleaved", IL): a row of up to 128 cells is 4 words; word w holds the columns c with c % 4 == w, bit b of
// word w <-> column 4b + w.  The 4 columns 4b .. 4b+3 are therefore bit b of words 0..3, and a 32x32 bit-matrix transpose of
// (8 planes x 4 words) yields, per b, one 32-bit word with the 4 cells' bytes in column order.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define FFMP_HD __host__ __device__ __forceinline__
#else
#define FFMP_HD inline
#endif

namespace ffmp {
namespace rowops {

// PRMT with the sign-replication mode (selector nibble bit 3): byte i of the result = byte (s_i & 7) of {b:a}, or 0x00 / 0xFF
// by that byte's msb when s_i & 8.
FFMP_HD uint32_t prmt(uint32_t a, uint32_t b, uint32_t s) {
#if defined(__CUDA_ARCH__)
    uint32_t r;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(s));
    return r;
#else
    const uint64_t t = (static_cast<uint64_t>(b) << 32) | a;
    uint32_t r = 0;
    for (int i = 0; i < 4; ++i) {
        const uint32_t sel = (s >> (4 * i)) & 0xFu;
        uint32_t byte = static_cast<uint32_t>(t >> (8 * (sel & 7u))) & 0xFFu;
        if (sel & 8u) byte = (byte & 0x80u) ? 0xFFu : 0u;
        r |= byte << (8 * i);
    }
    return r;
#endif
}

// One block-swap stage of a bit-matrix transpose on the word pair (a, b) = (x[m], x[m + J]): the J-wide column blocks
// (a: high, b: low) are exchanged.  J = 16 / 8 are byte moves (PRMT), J = 4 / 2 / 1 are bit selects (shift + LOP3).
template <int J>
FFMP_HD void swap_stage(uint32_t &a, uint32_t &b) {
    if constexpr (J == 16) {
        const uint32_t na = prmt(a, b, 0x5410u), nb = prmt(a, b, 0x7632u);
        a = na; b = nb;
    } else if constexpr (J == 8) {
        const uint32_t na = prmt(a, b, 0x6240u), nb = prmt(a, b, 0x7351u);
        a = na; b = nb;
    } else {
        constexpr uint32_t M = J == 4 ? 0x0F0F0F0Fu : J == 2 ? 0x33333333u : 0x55555555u;
        const uint32_t na = (a & M) | ((b << J) & ~M);
        const uint32_t nb = ((a >> J) & M) | (b & ~M);
        a = na; b = nb;
    }
}

template <int NW, int J>
FFMP_HD void transpose_stage(uint32_t (&x)[NW]) {
#pragma unroll
    for (int m = 0; m < NW; ++m)
        if ((m & J) == 0 && m + J < NW) swap_stage<J>(x[m], x[m + J]);
}

// 32 x 32 bit-matrix transpose in place: out[b] bit m = in[m] bit b.
FFMP_HD void transpose32(uint32_t (&x)[32]) {
    transpose_stage<32, 16>(x); transpose_stage<32, 8>(x); transpose_stage<32, 4>(x);
    transpose_stage<32, 2>(x); transpose_stage<32, 1>(x);
}

// Two 16 x 16 bit-matrix transposes side by side (the low and the high halves of 16 words):
// out[m] bit q = in[q] bit m and out[m] bit 16 + q = in[q] bit 16 + m (q, m < 16).
FFMP_HD void transpose16x2(uint32_t (&x)[16]) {
    transpose_stage<16, 8>(x); transpose_stage<16, 4>(x); transpose_stage<16, 2>(x); transpose_stage<16, 1>(x);
}

// ---- input: 16 occupancy bytes -> occupied bits ------------------------------------------------------------------------
// x = the bytes of columns 16q .. 16q+15 of a row (x.i = columns 16q + 4i .. + 3).  Result: byte w = the 4 occupied bits
// (byte != 0) of the cells b = 4q .. 4q+3 of interleaved word w, in its low nibble.
FFMP_HD uint32_t nonzero_msb(uint32_t x) { return (((x & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | x) & 0x80808080u; }
FFMP_HD uint32_t occupied_nibbles(uint32_t x0, uint32_t x1, uint32_t x2, uint32_t x3) {
    return (nonzero_msb(x0) >> 7) | (nonzero_msb(x1) >> 6) | (nonzero_msb(x2) >> 5) | (nonzero_msb(x3) >> 4);
}

// ---- flow direction (SPEC.md §4 F2), bit-parallel on one interleaved row ---------------------------------------------------
// Neighbour columns in the interleaved layout: the cells at column - 1 / column + 1 of word w's cells are word w -/+ 1; only
// the wrap-around words need a shift.
FFMP_HD uint32_t il_lo(const uint32_t (&x)[4], int w) { return w > 0 ? x[w - 1] : x[3] << 1; }   // bit b <- column - 1
FFMP_HD uint32_t il_hi(const uint32_t (&x)[4], int w) { return w < 3 ? x[w + 1] : x[0] >> 1; }   // bit b <- column + 1

// One row R of the grid.  c = row R, u = row R-1 ("west", code 4), d = row R+1 ("east", code 0); index i <-> x, j <-> y as in
// SPEC.md §2, so N / S are column + 1 / column - 1.  b1* / b2* = cost bits 1 / 2, V* = reached, F* = free; par0 = cost bit 0 of
// the row's cells of word 0 (bit 0 of a 4-connected unit-cost BFS distance is the checkerboard colour: word w has par0 ^ (w&1)).
// Output planes: n[0..2] = table index bits, n[3] = sign-mode bit of the PRMT selector nibble of flow_lookup().
struct RowIn {
    uint32_t b1c[4], b2c[4], Vc[4], Fc[4];
    uint32_t b1u[4], b2u[4], Vu[4], Fu[4];
    uint32_t b1d[4], b2d[4], Vd[4], Fd[4];
};

FFMP_HD void direction_nibbles(const RowIn &in, uint32_t par0, uint32_t (&n)[4][4]) {
#pragma unroll
    for (int w = 0; w < 4; ++w) {
        const uint32_t b0 = ((par0 ^ static_cast<uint32_t>(w)) & 1u) ? 0xFFFFFFFFu : 0u;
        const uint32_t own = in.Vc[w];
        const uint32_t t = in.b1c[w] ^ ~b0;          // bit 1 of (cost - 1)
        const uint32_t u = in.b2c[w] ^ ~in.b1c[w];   // bit 2 of (cost - 2)
        // orthogonal neighbours one level lower (codes 0 E, 2 N, 4 W, 6 S)
        const uint32_t lE = own & in.Vd[w] & ~(in.b1d[w] ^ t);
        const uint32_t lW = own & in.Vu[w] & ~(in.b1u[w] ^ t);
        const uint32_t lN = own & il_hi(in.Vc, w) & ~(il_hi(in.b1c, w) ^ t);
        const uint32_t lS = own & il_lo(in.Vc, w) & ~(il_lo(in.b1c, w) ^ t);
        // admissible diagonals two levels lower (codes 1 NE, 3 NW, 5 SW, 7 SE): both side cells free
        const uint32_t fE = in.Fd[w], fW = in.Fu[w], fN = il_hi(in.Fc, w), fS = il_lo(in.Fc, w);
        const uint32_t lNE = own & il_hi(in.Fd, w) & fE & fN & (il_hi(in.b1d, w) ^ in.b1c[w]) & ~(il_hi(in.b2d, w) ^ u);
        const uint32_t lNW = own & il_hi(in.Fu, w) & fW & fN & (il_hi(in.b1u, w) ^ in.b1c[w]) & ~(il_hi(in.b2u, w) ^ u);
        const uint32_t lSW = own & il_lo(in.Fu, w) & fW & fS & (il_lo(in.b1u, w) ^ in.b1c[w]) & ~(il_lo(in.b2u, w) ^ u);
        const uint32_t lSE = own & il_lo(in.Fd, w) & fE & fS & (il_lo(in.b1d, w) ^ in.b1c[w]) & ~(il_lo(in.b2d, w) ^ u);
        // first minimum in the scan order E, NE, N, NW, W, SW, S, SE with strict '<': a diagonal (cost - 2) beats every
        // orthogonal neighbour (cost - 1); among equals the lowest code wins
        const uint32_t anyD = lNE | lNW | lSW | lSE;
        const uint32_t m0 = (anyD & lNE) | (~anyD & lE);
        const uint32_t m1 = (anyD & lNW) | (~anyD & lN);
        const uint32_t m2 = (anyD & lSW) | (~anyD & lW);
        const uint32_t m3 = (anyD & lSE) | (~anyD & lS);
        const uint32_t d1 = ~m0 & (m1 | (~m2 & m3));      // code = (d2 d1 anyD), none = no lower neighbour
        const uint32_t d2 = ~m0 & ~m1 & (m2 | m3);
        const uint32_t some = m0 | m1 | m2 | m3;
        // selector nibble of flow_lookup(): code k = 1..7 -> k; code 0 -> sign mode on entry 1 (0x1C -> 0x00);
        // none (goal, unreached) -> entry 0 (224); occupied -> sign mode on entry 0 (0xE0 -> 0xFF)
        const uint32_t zero = some & ~(anyD | d1 | d2);   // a lower neighbour exists and it is E
        n[0][w] = anyD | zero;
        n[1][w] = d1;
        n[2][w] = d2;
        n[3][w] = ~in.Fc[w] | zero;
    }
}

// the four flow bytes (SPEC.md §5: 255 occupied, else code * 28, 224 = none) of one selector (low 16 bits of sel)
FFMP_HD uint32_t flow_lookup(uint32_t sel) { return prmt(0x54381CE0u, 0xC4A88C70u, sel); }

// n[q][w] (q = nibble bit, w = word) -> the row's 32 flow words (word b = columns 4b .. 4b+3)
FFMP_HD void flow_row_words(const uint32_t (&n)[4][4], uint32_t (&out)[32]) {
    uint32_t x[16];
#pragma unroll
    for (int w = 0; w < 4; ++w)
#pragma unroll
        for (int q = 0; q < 4; ++q) x[4 * w + q] = n[q][w];
    transpose16x2(x);      // x[m]: low half = the selector of b = m, high half = the selector of b = 16 + m
#pragma unroll
    for (int m = 0; m < 16; ++m) {
        out[m] = flow_lookup(x[m]);
        out[16 + m] = flow_lookup(x[m] >> 16);
    }
}

// ---- integration field (SPEC.md §4 F1) of one row as int32, depth < 256 ---------------------------------------------------
// bin[k][w] = cost bit k + 1 (k = 0..6), V = reached, par0 as above.  out[b] = the int4 of columns 4b .. 4b+3
// (0x7FFFFFFF where not reached).
struct Int4 { int32_t x, y, z, w; };

// T = one word of the transposed planes: byte w = (cost & 0xFE) | not-reached of column 4b + w (0xFF where not reached);
// par = cost bit 0 of the four columns.  One PRMT per cell: (cost byte, 0, 0, 0) or 0x7FFFFFFF.
FFMP_HD Int4 widen_cost4(uint32_t T, uint32_t par) {
    const uint32_t cst = T | par;                           // reached: the cost byte; else 0xFF
    const uint32_t m4 = prmt(T << 7, 0u, 0xba98u);          // 0xFF per byte where not reached
    const uint32_t h4 = m4 & 0x7F7F7F7Fu;
    const uint32_t lo = prmt(cst, m4, 0x5410u), hi = prmt(cst, m4, 0x7632u);   // (c0 c1 m0 m1) / (c2 c3 m2 m3)
    Int4 o;
    o.x = static_cast<int32_t>(prmt(lo, h4, 0x4220u));
    o.y = static_cast<int32_t>(prmt(lo, h4, 0x5331u));
    o.z = static_cast<int32_t>(prmt(hi, h4, 0x6220u));
    o.w = static_cast<int32_t>(prmt(hi, h4, 0x7331u));
    return o;
}

FFMP_HD void cost_row_words(const uint32_t (&bin)[7][4], const uint32_t (&V)[4], uint32_t par0, Int4 (&out)[32]) {
    uint32_t x[32];
#pragma unroll
    for (int w = 0; w < 4; ++w) {
        x[8 * w] = ~V[w];                                      // bit 0 of the byte: not reached
#pragma unroll
        for (int k = 0; k < 7; ++k) x[8 * w + 1 + k] = bin[k][w] | ~V[w];   // bits 1..7: cost bits, all ones where not reached
    }
    transpose32(x);        // x[b] byte w = (cost & 0xFE) | not-reached of column 4b + w; 0xFF where not reached
    const uint32_t par = (par0 & 1u) ? 0x00010001u : 0x01000100u;   // cost bit 0 of the four columns
#pragma unroll
    for (int b = 0; b < 32; ++b) out[b] = widen_cost4(x[b], par);
}


// ---- wide rows (large maps): WPR interleaved words per row, word w = the columns c % WPR == w, bit b <-> column b * WPR + w ----
// The four columns b * WPR + 4j .. + 3 are bit b of words 4j .. 4j+3, so every group j of four words goes through the same
// transposes as a 128-column row and yields, per b, the 4 cells at columns b * WPR + 4j.
template <int WPR> FFMP_HD uint32_t ilw_lo(const uint32_t (&x)[WPR], int w) { return w > 0 ? x[w - 1] : x[WPR - 1] << 1; }
template <int WPR> FFMP_HD uint32_t ilw_hi(const uint32_t (&x)[WPR], int w) { return w < WPR - 1 ? x[w + 1] : x[0] >> 1; }

template <int WPR>
struct RowInW {
    uint32_t b1c[WPR], b2c[WPR], Vc[WPR], Fc[WPR];
    uint32_t b1u[WPR], b2u[WPR], Vu[WPR];      // row R-1 (a free neighbour of a reached cell is reached: V stands in for F)
    uint32_t b1d[WPR], b2d[WPR], Vd[WPR];      // row R+1
};

// direction_nibbles for a wide row; n[q][w] as above (WPR even: cost bit 0 of word w's cells is par0 ^ (w & 1))
template <int WPR>
FFMP_HD void direction_nibbles_w(const RowInW<WPR> &in, uint32_t par0, uint32_t (&n)[4][WPR]) {
#pragma unroll
    for (int w = 0; w < WPR; ++w) {
        const uint32_t b0 = ((par0 ^ static_cast<uint32_t>(w)) & 1u) ? 0xFFFFFFFFu : 0u;
        const uint32_t own = in.Vc[w];
        const uint32_t t = in.b1c[w] ^ ~b0;
        const uint32_t u = in.b2c[w] ^ ~in.b1c[w];
        const uint32_t lE = own & in.Vd[w] & ~(in.b1d[w] ^ t);
        const uint32_t lW = own & in.Vu[w] & ~(in.b1u[w] ^ t);
        const uint32_t lN = own & ilw_hi<WPR>(in.Vc, w) & ~(ilw_hi<WPR>(in.b1c, w) ^ t);
        const uint32_t lS = own & ilw_lo<WPR>(in.Vc, w) & ~(ilw_lo<WPR>(in.b1c, w) ^ t);
        const uint32_t fE = in.Vd[w], fW = in.Vu[w], fN = ilw_hi<WPR>(in.Vc, w), fS = ilw_lo<WPR>(in.Vc, w);
        const uint32_t lNE = own & ilw_hi<WPR>(in.Vd, w) & fE & fN & (ilw_hi<WPR>(in.b1d, w) ^ in.b1c[w]) & ~(ilw_hi<WPR>(in.b2d, w) ^ u);
        const uint32_t lNW = own & ilw_hi<WPR>(in.Vu, w) & fW & fN & (ilw_hi<WPR>(in.b1u, w) ^ in.b1c[w]) & ~(ilw_hi<WPR>(in.b2u, w) ^ u);
        const uint32_t lSW = own & ilw_lo<WPR>(in.Vu, w) & fW & fS & (ilw_lo<WPR>(in.b1u, w) ^ in.b1c[w]) & ~(ilw_lo<WPR>(in.b2u, w) ^ u);
        const uint32_t lSE = own & ilw_lo<WPR>(in.Vd, w) & fE & fS & (ilw_lo<WPR>(in.b1d, w) ^ in.b1c[w]) & ~(ilw_lo<WPR>(in.b2d, w) ^ u);
        const uint32_t anyD = lNE | lNW | lSW | lSE;
        const uint32_t m0 = (anyD & lNE) | (~anyD & lE);
        const uint32_t m1 = (anyD & lNW) | (~anyD & lN);
        const uint32_t m2 = (anyD & lSW) | (~anyD & lW);
        const uint32_t m3 = (anyD & lSE) | (~anyD & lS);
        const uint32_t d1 = ~m0 & (m1 | (~m2 & m3));
        const uint32_t d2 = ~m0 & ~m1 & (m2 | m3);
        const uint32_t some = m0 | m1 | m2 | m3;
        const uint32_t zero = some & ~(anyD | d1 | d2);
        n[0][w] = anyD | zero;
        n[1][w] = d1;
        n[2][w] = d2;
        n[3][w] = ~in.Fc[w] | zero;
    }
}

// the flow words of word group j of a wide row: out[b] = the 4 flow bytes at columns b * WPR + 4j .. + 3
template <int WPR>
FFMP_HD void flow_group_words(const uint32_t (&n)[4][WPR], int j, uint32_t (&out)[32]) {
    uint32_t x[16];
#pragma unroll
    for (int wl = 0; wl < 4; ++wl)
#pragma unroll
        for (int q = 0; q < 4; ++q) x[4 * wl + q] = n[q][4 * j + wl];
    transpose16x2(x);
#pragma unroll
    for (int m = 0; m < 16; ++m) {
        out[m] = flow_lookup(x[m]);
        out[16 + m] = flow_lookup(x[m] >> 16);
    }
}

// T_lo / T_hi = transposed plane words of 4 cells: T_lo byte w = (cost & 0xFE) | not-reached (0xFF where not reached), T_hi byte w =
// cost bits 8..15 (anything where not reached).  -> the four int32 costs, 0x7FFFFFFF where not reached.
FFMP_HD Int4 widen_cost4_16(uint32_t T_lo, uint32_t T_hi, uint32_t par) {
    const uint32_t cst = T_lo | par;
    const uint32_t m4 = prmt(T_lo << 7, 0u, 0xba98u);           // 0xFF per byte where not reached
    const uint32_t hi = T_hi | m4;                              // second byte: cost bits 8..15, 0xFF where not reached
    const uint32_t h4 = m4 & 0x7F7F7F7Fu;
    const uint32_t a01 = prmt(cst, hi, 0x5140u), a23 = prmt(cst, hi, 0x7362u);      // (c0 h0 c1 h1) / (c2 h2 c3 h3)
    const uint32_t mh = prmt(m4, h4, 0x5410u);                  // (m0 m1 x0 x1): third byte source / top byte source for cells 0, 1
    const uint32_t mh2 = prmt(m4, h4, 0x7632u);                 // ... for cells 2, 3
    Int4 o;
    o.x = static_cast<int32_t>(prmt(a01, mh, 0x6410u));          // (c0, h0, m0, x0)
    o.y = static_cast<int32_t>(prmt(a01, mh, 0x7532u));          // (c1, h1, m1, x1)
    o.z = static_cast<int32_t>(prmt(a23, mh2, 0x6410u));
    o.w = static_cast<int32_t>(prmt(a23, mh2, 0x7532u));
    return o;
}

// bytes of 16 columns 16q .. 16q+15 of a 512-wide row -> 16 occupied flags (bit w = byte w != 0): with WPR = 16 byte w of
// chunk q is bit q of interleaved word w, so the row's words are the 32 x 16 bit transpose of its 32 flag halfwords
FFMP_HD uint32_t occupied_flags16(uint32_t x0, uint32_t x1, uint32_t x2, uint32_t x3) {
    // msb of every non-zero byte -> one bit per byte, gathered with a multiply (bits 7, 15, 23, 31 -> a nibble)
    const uint32_t f0 = ((nonzero_msb(x0) >> 7) * 0x00204081u) >> 21 & 0xFu;
    const uint32_t f1 = ((nonzero_msb(x1) >> 7) * 0x00204081u) >> 21 & 0xFu;
    const uint32_t f2 = ((nonzero_msb(x2) >> 7) * 0x00204081u) >> 21 & 0xFu;
    const uint32_t f3 = ((nonzero_msb(x3) >> 7) * 0x00204081u) >> 21 & 0xFu;
    return f0 | (f1 << 4) | (f2 << 8) | (f3 << 12);
}

}  // namespace rowops
}  // namespace ffmp
