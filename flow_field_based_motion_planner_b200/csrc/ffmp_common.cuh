// ffmp_common.cuh — constants, hash RNG and fp32 elementary functions shared by the sm_100a kernels.
// Every function here implements one block of SPEC.md; fp32 code uses __f*_rn intrinsics so that no
// FMA contraction can occur regardless of compiler flags (one rounding per written operation).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace ffmp {

constexpr float RES = 0.05f;
constexpr float INV_RES = 20.0f;
constexpr float PI_F = 3.14159274f;
constexpr float TWO_PI_F = 6.28318548f;
constexpr float PI_2_F = 1.57079637f;
constexpr float PI_4_F = 0.785398185f;
constexpr int COST_INF = 0x7FFFFFFF;
constexpr uint32_t FULL = 0xFFFFFFFFu;

constexpr uint32_t S_START = 0x53544152u;
constexpr uint32_t S_GOAL = 0x474F414Cu;
constexpr uint32_t S_YAW = 0x59415721u;

// scenario record / state record word indices (include/ffmp_b200.h)
enum { SC_X0 = 0, SC_Y0, SC_YAW0, SC_GX, SC_GY, SC_GI, SC_GJ, SC_KEY, SC_WORDS };
enum { ST_X = 0, ST_Y, ST_YAW, ST_GX, ST_GY, ST_DFIRST, ST_RETURN, ST_STEPS, ST_EPISODE, ST_WORDS = 16 };

// ---- SPEC.md §3 hash ---------------------------------------------------------------------------
__host__ __device__ __forceinline__ uint32_t mix32(uint32_t x) {
    x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
    return x;
}
__host__ __device__ __forceinline__ uint32_t scenario_key(uint64_t seed, uint32_t env_gid, uint32_t episode) {
    uint32_t k = mix32(static_cast<uint32_t>(seed) ^ 0x9E3779B9u);
    k = mix32(k ^ static_cast<uint32_t>(seed >> 32));
    k = mix32(k ^ env_gid);
    k = mix32(k ^ episode);
    return k;
}
__host__ __device__ __forceinline__ uint32_t draw(uint32_t key, uint32_t stream, uint32_t t) {
    return mix32(mix32(key ^ stream) + t * 0x9E3779B1u);
}

// ---- SPEC.md §6 fp32 elementary functions --------------------------------------------------------
__device__ __forceinline__ float fmul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float fadd(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float fsub(float a, float b) { return __fsub_rn(a, b); }

__device__ __forceinline__ float pi_to_pi(float a) {
    while (a >= PI_F) a = fsub(a, TWO_PI_F);
    while (a <= -PI_F) a = fadd(a, TWO_PI_F);
    return a;
}

__device__ __forceinline__ void sincos_spec(float a, float &s, float &c) {
    const float q = rintf(fmul(a, 0.636619747f));
    float r = fsub(a, fmul(q, 1.5703125f));
    r = fsub(r, fmul(q, 4.837512969970703125e-4f));
    r = fsub(r, fmul(q, 7.54978995489188216e-8f));
    const float z = fmul(r, r);
    float sp = fmul(-1.9515295891e-4f, z);
    sp = fadd(sp, 8.3321608736e-3f);
    sp = fmul(sp, z);
    sp = fsub(sp, 1.6666654611e-1f);
    sp = fmul(sp, z);
    sp = fmul(sp, r);
    sp = fadd(sp, r);
    float cp = fmul(2.443315711809948e-5f, z);
    cp = fsub(cp, 1.388731625493765e-3f);
    cp = fmul(cp, z);
    cp = fadd(cp, 4.166664568298827e-2f);
    cp = fmul(cp, z);
    cp = fmul(cp, z);
    cp = fsub(cp, fmul(0.5f, z));
    cp = fadd(cp, 1.0f);
    const int n = static_cast<int>(q) & 3;
    if (n == 0) { s = sp; c = cp; }
    else if (n == 1) { s = cp; c = -sp; }
    else if (n == 2) { s = -sp; c = -cp; }
    else { s = -cp; c = sp; }
}

__device__ __forceinline__ float atan2_spec(float y, float x) {
    if (x == 0.0f && y == 0.0f) return 0.0f;
    const float ax = fabsf(x), ay = fabsf(y);
    const bool swap = ay > ax;
    const float t = __fdiv_rn(swap ? ax : ay, swap ? ay : ax);
    float base, u;
    if (t > 0.414213568f) { base = PI_4_F; u = __fdiv_rn(fsub(t, 1.0f), fadd(t, 1.0f)); }
    else { base = 0.0f; u = t; }
    const float z = fmul(u, u);
    float p = fmul(8.05374449538e-2f, z);
    p = fsub(p, 1.38776856032e-1f);
    p = fmul(p, z);
    p = fadd(p, 1.99777106478e-1f);
    p = fmul(p, z);
    p = fsub(p, 3.33329491539e-1f);
    p = fmul(p, z);
    p = fmul(p, u);
    p = fadd(p, u);
    float a = fadd(base, p);
    if (swap) a = fsub(PI_2_F, a);
    if (x < 0.0f) a = fsub(PI_F, a);
    if (y < 0.0f) a = -a;
    return a;
}

__device__ __forceinline__ float dist_spec(float dx, float dy) {
    return __fsqrt_rn(fadd(fmul(dx, dx), fmul(dy, dy)));
}

__device__ __forceinline__ int robot_cell(float x) {
    return static_cast<int>(floorf(fadd(fmul(x, INV_RES), 0.5f)));
}

// ---- SPEC.md §3 scenario sampling (shared by scenario.cu and the flow-field kernels) -------------
__device__ __forceinline__ void cell_of(uint32_t u, int G, int &i, int &j) {
    const uint32_t span = static_cast<uint32_t>(G - 6);
    i = 3 + static_cast<int>((u & 0xFFFFu) % span);
    j = 3 + static_cast<int>((u >> 16) % span);
}

struct ScenarioParams {
    int si, sj, gi, gj;
    float yaw;
};

__device__ __forceinline__ ScenarioParams sample_scenario(uint32_t key, int G, int goal_mode) {
    ScenarioParams p;
    if (goal_mode == 0) {
        cell_of(draw(key, S_START, 0), G, p.si, p.sj);
        for (uint32_t t = 0; t < 64; ++t) {
            cell_of(draw(key, S_GOAL, t), G, p.gi, p.gj);
            if ((p.gi - p.si) * (p.gi - p.si) + (p.gj - p.sj) * (p.gj - p.sj) >= 400) break;
        }
    } else {
        p.gi = G - 8; p.gj = G - 8;
        for (uint32_t t = 0; t < 64; ++t) {
            cell_of(draw(key, S_START, t), G, p.si, p.sj);
            if ((p.gi - p.si) * (p.gi - p.si) + (p.gj - p.sj) * (p.gj - p.sj) >= 400) break;
        }
    }
    p.yaw = pi_to_pi(fsub(fmul(static_cast<float>(draw(key, S_YAW, 0) >> 8), TWO_PI_F * 5.9604644775390625e-08f), PI_F));
    return p;
}

__device__ __forceinline__ void store_scenario_record(uint32_t *rec, const ScenarioParams &p, uint32_t key) {
    *reinterpret_cast<uint4 *>(rec) = make_uint4(__float_as_uint(fmul(static_cast<float>(p.si), RES)),
                                                __float_as_uint(fmul(static_cast<float>(p.sj), RES)), __float_as_uint(p.yaw),
                                                __float_as_uint(fmul(static_cast<float>(p.gi), RES)));
    *reinterpret_cast<uint4 *>(rec + 4) = make_uint4(__float_as_uint(fmul(static_cast<float>(p.gj), RES)),
                                                    static_cast<uint32_t>(p.gi), static_cast<uint32_t>(p.gj), key);
}

// bits [lo, hi] (inclusive, clipped to the 32-column word starting at column c0) as a mask
__device__ __forceinline__ uint32_t col_range_mask(int lo, int hi, int c0) {
    lo = max(lo - c0, 0); hi = min(hi - c0, 31);
    if (lo > hi) return 0u;
    return (0xFFFFFFFFu >> (31 - hi)) & (0xFFFFFFFFu << lo);
}

// free-cell mask of the 32 cells (row R, columns c0..c0+31) of a generated scenario (bit = 1: free)
__device__ __forceinline__ uint32_t scenario_free_word(uint32_t key, int R, int c0, int G, int bs, uint32_t p_thresh,
                                                       const ScenarioParams &p) {
    if (R <= 0 || R >= G - 1) return 0u;                                    // border rows / padding rows
    const int bw = 1 << bs;
    uint32_t occ = 0;
    if (bw >= 32) {
        const uint32_t blk = (static_cast<uint32_t>(R >> bs) << 16) | static_cast<uint32_t>(c0 >> bs);
        occ = mix32(key + blk * 0x9E3779B1u) < p_thresh ? 0xFFFFFFFFu : 0u;
    } else {
        const uint32_t bmask = (1u << bw) - 1u;
        for (int c = 0; c < 32; c += bw) {
            const uint32_t blk = (static_cast<uint32_t>(R >> bs) << 16) | static_cast<uint32_t>((c0 + c) >> bs);
            if (mix32(key + blk * 0x9E3779B1u) < p_thresh) occ |= bmask << c;
        }
    }
    uint32_t fr = ~occ;
    if (abs(R - p.si) <= 2) fr |= col_range_mask(p.sj - 2, p.sj + 2, c0);   // cleared 5x5 around the start
    if (abs(R - p.gi) <= 2) fr |= col_range_mask(p.gj - 2, p.gj + 2, c0);   // ... and around the goal
    return fr & col_range_mask(1, G - 2, c0);                               // border columns / padding
}

// Column-interleaved form of the same row (rows of up to 128 cells as 4 words: word w bit b <-> column 4b + w), used by the
// G <= 128 flow-field kernel.  Columns lo..hi (inclusive) as a mask of word w:
__device__ __forceinline__ uint32_t il_col_range_mask(int lo, int hi, int w) {
    int blo = (lo - w + 3) >> 2, bhi = (hi - w) >> 2;      // ceil / floor of (col - w) / 4 (arithmetic shifts)
    blo = max(blo, 0); bhi = min(bhi, 31);
    if (blo > bhi) return 0u;
    return (0xFFFFFFFFu >> (31 - bhi)) & (0xFFFFFFFFu << blo);
}

__device__ __forceinline__ void scenario_free_row_il(uint32_t key, int R, int G, int bs, uint32_t p_thresh,
                                                     const ScenarioParams &p, uint32_t (&out)[4]) {
    if (R <= 0 || R >= G - 1) {                                             // border rows / padding rows
        out[0] = out[1] = out[2] = out[3] = 0u;
        return;
    }
    const uint32_t rblk = static_cast<uint32_t>(R >> bs) << 16;
    uint32_t occ[4];
    if (bs >= 2) {
        // the block column of column 4b + w is b >> (bs - 2) for every w: one occupancy word serves the four words
        const int sb = bs - 2;
        uint32_t o = 0;
        if (sb >= 5) {
            o = mix32(key + rblk * 0x9E3779B1u) < p_thresh ? 0xFFFFFFFFu : 0u;
        } else {
            const int bw = 1 << sb;
            const uint32_t bmask = (1u << bw) - 1u;
            for (int b = 0; b < 32; b += bw)
                if (mix32(key + (rblk | static_cast<uint32_t>(b >> sb)) * 0x9E3779B1u) < p_thresh) o |= bmask << b;
        }
        occ[0] = occ[1] = occ[2] = occ[3] = o;
    } else {
#pragma unroll
        for (int w = 0; w < 4; ++w) {
            uint32_t o = 0;
            for (int b = 0; b < 32; ++b)
                if (mix32(key + (rblk | static_cast<uint32_t>((4 * b + w) >> bs)) * 0x9E3779B1u) < p_thresh) o |= 1u << b;
            occ[w] = o;
        }
    }
#pragma unroll
    for (int w = 0; w < 4; ++w) {
        uint32_t fr = ~occ[w];
        if (abs(R - p.si) <= 2) fr |= il_col_range_mask(p.sj - 2, p.sj + 2, w);   // cleared 5x5 around the start
        if (abs(R - p.gi) <= 2) fr |= il_col_range_mask(p.gj - 2, p.gj + 2, w);   // ... and around the goal
        out[w] = fr & il_col_range_mask(1, G - 2, w);                             // border columns / padding
    }
}

// SPEC.md §1 action table (robot/config.py:25-58)
__device__ __forceinline__ void action_lookup(int a, float &v, float &w) {
    const int iv = a / 7, iw = a - 7 * iv;
    v = iv == 0 ? 0.0f : iv == 1 ? 0.2f : iv == 2 ? 0.4f : 0.6f;
    w = iw == 0 ? -0.6f : iw == 1 ? -0.4f : iw == 2 ? -0.2f : iw == 3 ? 0.0f : iw == 4 ? 0.2f : iw == 5 ? 0.4f : 0.6f;
}

}  // namespace ffmp
