// Host build of csrc/flow_rowops.cuh (g++, no GPU): the per-row bit-plane arithmetic of the interleaved flow-field kernel,
// driven by a sequential restatement of the kernel's glue (same data layout, same Gray-code-by-addition level record, same
// checkerboard rule for cost bit 0).  tests/test_rowops_host.py compares the result with the CPU oracle.
// TEST INFRASTRUCTURE: nothing here is linked into libffmp_b200.so.
#include <cstdint>
#include <cstring>
#include <vector>

#include "../../flow_field_based_motion_planner_b200/csrc/flow_rowops.cuh"

using namespace ffmp::rowops;

namespace {
constexpr int P = 128;
constexpr int NPL = 18;
constexpr int32_t COST_INF = 0x7FFFFFFF;
struct Plane { uint32_t w[P][4]; };
}  // namespace

extern "C" {

// out[b] bit m = in[m] bit b
void host_transpose32(const uint32_t *in, uint32_t *out) {
    uint32_t x[32];
    std::memcpy(x, in, sizeof(x));
    transpose32(x);
    std::memcpy(out, x, sizeof(x));
}

void host_transpose16x2(const uint32_t *in, uint32_t *out) {
    uint32_t x[16];
    std::memcpy(x, in, sizeof(x));
    transpose16x2(x);
    std::memcpy(out, x, sizeof(x));
}

uint32_t host_prmt(uint32_t a, uint32_t b, uint32_t s) { return prmt(a, b, s); }
uint32_t host_occupied_nibbles(uint32_t x0, uint32_t x1, uint32_t x2, uint32_t x3) { return occupied_nibbles(x0, x1, x2, x3); }

// The whole interleaved kernel for one grid (96 < G <= 128).  Returns the number of BFS levels executed.
int il_emulate_grid(const uint8_t *occ, int G, int gi, int gj, int32_t *cost, uint8_t *flow) {
    static Plane fr, A, F, V, g[NPL], bin[NPL];
    std::memset(&fr, 0, sizeof(fr));
    for (int k = 0; k < NPL; ++k) std::memset(&g[k], 0, sizeof(Plane));
    // ---- input: bytes -> interleaved free mask (the 16-byte path of the kernel when G == 128) ----
    for (int R = 0; R < G; ++R) {
        if (G == 128) {
            uint32_t o[4] = {0, 0, 0, 0};
            for (int q = 0; q < 8; ++q) {
                uint32_t x[4];
                std::memcpy(x, occ + R * G + 16 * q, 16);
                const uint32_t z = occupied_nibbles(x[0], x[1], x[2], x[3]);
                for (int w = 0; w < 4; ++w) o[w] += ((z >> (8 * w)) & 0xFu) << (4 * q);
            }
            for (int w = 0; w < 4; ++w) fr.w[R][w] = ~o[w];
        } else {
            for (int c = 0; c < G; ++c)
                if (occ[R * G + c] == 0) fr.w[R][c & 3] |= 1u << (c >> 2);
        }
    }
    A = fr;
    std::memset(&F, 0, sizeof(F));
    if (gi >= 0 && gj >= 0 && gi < G && gj < G) {
        const uint32_t bit = (1u << (gj >> 2)) & A.w[gi][gj & 3];
        F.w[gi][gj & 3] = bit;
        A.w[gi][gj & 3] ^= bit;
    }
    auto step = [&]() {
        Plane N;
        for (int R = 0; R < P; ++R)
            for (int w = 0; w < 4; ++w) {
                const uint32_t up = R > 0 ? F.w[R - 1][w] : 0u, dn = R < P - 1 ? F.w[R + 1][w] : 0u;
                N.w[R][w] = (il_lo(F.w[R], w) | il_hi(F.w[R], w) | up | dn) & A.w[R][w];
            }
        for (int R = 0; R < P; ++R)
            for (int w = 0; w < 4; ++w) { A.w[R][w] -= N.w[R][w]; F.w[R][w] = N.w[R][w]; }
    };
    auto gray = [&](uint32_t L) {        // L even: plane ctz(L >> 1) += +-avail (nested sets: the alternating sum is the XOR)
        const uint32_t M = L >> 1;
        const int k = __builtin_ctz(M);
        const uint32_t s = ((M >> (k + 1)) & 1u) ? 0xFFFFFFFFu : 1u;
        for (int R = 0; R < P; ++R)
            for (int w = 0; w < 4; ++w) g[k].w[R][w] += s * A.w[R][w];
    };
    uint32_t L = 1;
    for (;; L += 4) {
        for (uint32_t h = 0; h < 4; h += 2) { step(); gray(L + h + 1); step(); }
        uint32_t any = 0;
        for (int R = 0; R < P; ++R)
            for (int w = 0; w < 4; ++w) any |= F.w[R][w];
        if (!any) break;
    }
    const uint32_t Lmax = L + 2;
    const uint32_t Mmax = Lmax >> 1;
    const int kmax = 32 - __builtin_clz(Mmax);
    for (int R = 0; R < P; ++R)
        for (int w = 0; w < 4; ++w) V.w[R][w] = fr.w[R][w] & ~A.w[R][w];
    for (int k = 0; k < NPL; ++k) std::memset(&bin[k], 0, sizeof(Plane));
    for (int R = 0; R < P; ++R)
        for (int w = 0; w < 4; ++w) {
            uint32_t acc = 0;
            for (int k = kmax - 1; k >= 0; --k) { acc ^= g[k].w[R][w]; bin[k].w[R][w] = acc; }
        }
    // ---- per row: direction, flow bytes, cost ----
    const uint32_t zero4[4] = {0, 0, 0, 0};
    for (int R = 0; R < G; ++R) {
        RowIn in;
        auto put = [&](uint32_t (&dst)[4], const uint32_t *src) { std::memcpy(dst, src, 16); };
        put(in.b1c, bin[0].w[R]); put(in.b2c, bin[1].w[R]); put(in.Vc, V.w[R]); put(in.Fc, fr.w[R]);
        put(in.b1u, R > 0 ? bin[0].w[R - 1] : zero4); put(in.b2u, R > 0 ? bin[1].w[R - 1] : zero4);
        put(in.Vu, R > 0 ? V.w[R - 1] : zero4); put(in.Fu, in.Vu);   // a free neighbour of a reached cell is reached
        put(in.b1d, R < P - 1 ? bin[0].w[R + 1] : zero4); put(in.b2d, R < P - 1 ? bin[1].w[R + 1] : zero4);
        put(in.Vd, R < P - 1 ? V.w[R + 1] : zero4); put(in.Fd, in.Vd);
        const uint32_t par0 = static_cast<uint32_t>(R + gi + gj) & 1u;
        uint32_t n[4][4], fw[32];
        direction_nibbles(in, par0, n);
        flow_row_words(n, fw);
        for (int b = 0; 4 * b < G; ++b) std::memcpy(flow + R * G + 4 * b, &fw[b], 4);
        if (kmax <= 7) {
            uint32_t bk[7][4];
            for (int k = 0; k < 7; ++k) put(bk[k], bin[k].w[R]);
            Int4 c[32];
            cost_row_words(bk, in.Vc, par0, c);
            for (int b = 0; 4 * b < G; ++b) std::memcpy(cost + R * G + 4 * b, &c[b], 16);
        } else {
            for (int col = 0; col < G; ++col) {
                const int w = col & 3, b = col >> 2;
                int32_t v = COST_INF;
                if ((V.w[R][w] >> b) & 1u) {
                    v = static_cast<int32_t>((par0 ^ static_cast<uint32_t>(w)) & 1u);
                    for (int k = 0; k < kmax; ++k) v |= static_cast<int32_t>((bin[k].w[R][w] >> b) & 1u) << (k + 1);
                }
                cost[R * G + col] = v;
            }
        }
    }
    return static_cast<int>(Lmax);
}

}  // extern "C"

// ---- wide rows (flow_field_wide.cu): 384 < G <= 512, 16 interleaved words per row --------------------------------------------
namespace {
constexpr int PW = 512, WW = 16, NPLW = 20;
struct PlaneW { uint32_t w[PW][WW]; };
}  // namespace

extern "C" int ilw_emulate_grid(const uint8_t *occ, int G, int gi, int gj, int32_t *cost, uint8_t *flow) {
    static PlaneW fr, A, F, V, N, g[NPLW], bin[NPLW];
    std::memset(&fr, 0, sizeof(fr));
    for (int k = 0; k < NPLW; ++k) std::memset(&g[k], 0, sizeof(PlaneW));
    for (int R = 0; R < G; ++R) {
        if (G == 512) {
            // the kernel's path: 32 chunks of 16 bytes -> 32 flag halfwords -> 16 x 32 bit transpose
            uint32_t x[16];
            for (int q = 0; q < 16; ++q) {
                uint32_t a4[4], b4[4];
                std::memcpy(a4, occ + R * G + 16 * q, 16);
                std::memcpy(b4, occ + R * G + 16 * (q + 16), 16);
                x[q] = occupied_flags16(a4[0], a4[1], a4[2], a4[3]) | (occupied_flags16(b4[0], b4[1], b4[2], b4[3]) << 16);
            }
            transpose16x2(x);
            for (int w = 0; w < WW; ++w) fr.w[R][w] = ~x[w];
        } else {
            for (int c = 0; c < G; ++c)
                if (occ[R * G + c] == 0) fr.w[R][c % WW] |= 1u << (c / WW);
        }
    }
    A = fr;
    std::memset(&F, 0, sizeof(F));
    if (gi >= 0 && gj >= 0 && gi < G && gj < G) {
        const uint32_t bit = (1u << (gj / WW)) & A.w[gi][gj % WW];
        F.w[gi][gj % WW] = bit;
        A.w[gi][gj % WW] ^= bit;
    }
    auto step = [&]() {
        for (int R = 0; R < PW; ++R)
            for (int w = 0; w < WW; ++w) {
                const uint32_t up = R > 0 ? F.w[R - 1][w] : 0u, dn = R < PW - 1 ? F.w[R + 1][w] : 0u;
                N.w[R][w] = (ilw_lo<WW>(F.w[R], w) | ilw_hi<WW>(F.w[R], w) | up | dn) & A.w[R][w];
            }
        for (int R = 0; R < PW; ++R)
            for (int w = 0; w < WW; ++w) { A.w[R][w] -= N.w[R][w]; F.w[R][w] = N.w[R][w]; }
    };
    auto gray = [&](uint32_t L) {
        const uint32_t M = L >> 1;
        const int k = __builtin_ctz(M);
        const uint32_t s = ((M >> (k + 1)) & 1u) ? 0xFFFFFFFFu : 1u;
        for (int R = 0; R < PW; ++R)
            for (int w = 0; w < WW; ++w) g[k].w[R][w] += s * A.w[R][w];
    };
    uint32_t L = 1;
    for (;; L += 4) {
        for (uint32_t h = 0; h < 4; h += 2) { step(); gray(L + h + 1); step(); }
        uint32_t any = 0;
        for (int R = 0; R < PW; ++R)
            for (int w = 0; w < WW; ++w) any |= F.w[R][w];
        if (!any) break;
    }
    const uint32_t Mmax = (L + 2) >> 1;
    const int kmax = 32 - __builtin_clz(Mmax);
    for (int R = 0; R < PW; ++R)
        for (int w = 0; w < WW; ++w) {
            V.w[R][w] = fr.w[R][w] & ~A.w[R][w];
            uint32_t acc = 0;
            for (int k = NPLW - 1; k >= 0; --k) { if (k < kmax) acc ^= g[k].w[R][w]; bin[k].w[R][w] = k < kmax ? acc : 0u; }
        }
    const uint32_t zero[WW] = {0};
    for (int R = 0; R < G; ++R) {
        RowInW<WW> in;
        auto put = [&](uint32_t (&dst)[WW], const uint32_t *src) { std::memcpy(dst, src, 4 * WW); };
        put(in.b1c, bin[0].w[R]); put(in.b2c, bin[1].w[R]); put(in.Vc, V.w[R]); put(in.Fc, fr.w[R]);
        put(in.b1u, R > 0 ? bin[0].w[R - 1] : zero); put(in.b2u, R > 0 ? bin[1].w[R - 1] : zero); put(in.Vu, R > 0 ? V.w[R - 1] : zero);
        put(in.b1d, R < PW - 1 ? bin[0].w[R + 1] : zero); put(in.b2d, R < PW - 1 ? bin[1].w[R + 1] : zero); put(in.Vd, R < PW - 1 ? V.w[R + 1] : zero);
        const uint32_t par0 = static_cast<uint32_t>(R + gi + gj) & 1u;
        uint32_t n[4][WW];
        direction_nibbles_w<WW>(in, par0, n);
        for (int j = 0; j < 4; ++j) {
            uint32_t fw[32];
            flow_group_words<WW>(n, j, fw);
            for (int b = 0; b < 32; ++b)
                if (b * WW + 4 * j < G) std::memcpy(flow + R * G + b * WW + 4 * j, &fw[b], 4);
            const uint32_t par = ((par0 ^ static_cast<uint32_t>(4 * j)) & 1u) ? 0x00010001u : 0x01000100u;
            if (kmax <= 15) {
                uint32_t xl[32], xh[32];
                for (int wl = 0; wl < 4; ++wl) {
                    const int w = 4 * j + wl;
                    xl[8 * wl] = ~V.w[R][w];
                    for (int k = 0; k < 7; ++k) xl[8 * wl + 1 + k] = bin[k].w[R][w] | ~V.w[R][w];
                    for (int k = 0; k < 8; ++k) xh[8 * wl + k] = bin[7 + k].w[R][w];
                }
                transpose32(xl);
                transpose32(xh);
                for (int b = 0; b < 32; ++b)
                    if (b * WW + 4 * j < G) {
                        const Int4 c = widen_cost4_16(xl[b], xh[b], par);
                        std::memcpy(cost + R * G + b * WW + 4 * j, &c, 16);
                    }
            } else {
                for (int b = 0; b < 32; ++b)
                    for (int wl = 0; wl < 4; ++wl) {
                        const int w = 4 * j + wl, col = b * WW + w;
                        if (col >= G) continue;
                        int32_t v = COST_INF;
                        if ((V.w[R][w] >> b) & 1u) {
                            v = static_cast<int32_t>((par0 ^ static_cast<uint32_t>(w)) & 1u);
                            for (int k = 0; k < kmax; ++k) v |= static_cast<int32_t>((bin[k].w[R][w] >> b) & 1u) << (k + 1);
                        }
                        cost[R * G + col] = v;
                    }
            }
        }
    }
    return static_cast<int>(L + 2);
}
