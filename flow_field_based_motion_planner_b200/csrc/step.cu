// step.cu — SPEC.md §7/§8: the per-step kernels of the batched FFMP environment.
//
// dynamics_kernel — ONE THREAD PER ENVIRONMENT: fp32 unicycle integration (SPEC K), relative goal /
//   velocity (train.py:174-188), the 21-cell footprint collision test on the global flow image
//   (ffmp.py:85-105), goal test, reward, done, truncation (ffmp.py:120-164, train.py:607-608), the
//   auto-reset onto the next pre-generated scenario slot and the regeneration request; it leaves a
//   32-byte crop order per env.  Latency-bound (three dependent memory round trips), ~100 B/env.
// observe_kernel — one CTA per environment: copies the ego-centred W x W window of the flow image into
//   the observation frame ring with coalesced 32-bit stores (byte-realigned with a funnel shift, eight
//   rows of loads in flight per thread).  HBM-bound: W^2 read + W^2 written per env-step (+ one more
//   frame on ring wrap / reset).  Launched with programmatic dependent launch so that its CTAs are
//   resident and waiting (griddepcontrol.wait) when dynamics_kernel retires.
#include <cstring>
#include <mutex>
#include "ffmp_kernels.cuh"

namespace ffmp {

namespace {

__constant__ int8_t FOOT_DI[21] = {-2, -2, -2, -1, -1, -1, -1, -1, 0, 0, 0, 0, 0, 1, 1, 1, 1, 1, 2, 2, 2};
__constant__ int8_t FOOT_DJ[21] = {-1, 0, 1, -2, -1, 0, 1, 2, -2, -1, 0, 1, 2, -2, -1, 0, 1, 2, -1, 0, 1};

// crop order written by dynamics_kernel for observe_kernel (8 words per env)
enum { OB_CI = 0, OB_CJ, OB_PI, OB_PJ, OB_PLANE, OB_FLAGS, OB_WORDS = 8 };   // flags: bit0 active, bit1 two frames

// Per-lane plan for one 4-byte output word of a crop row.  The byte misalignment m of the window
// against the 4-byte aligned source row is the same for every word of a frame; out-of-grid columns
// are forced to 255 through `ormask`, and the two aligned source words are clamped into the row so
// that every load is in bounds (G % 4 == 0: an aligned word is either fully in or fully out).
struct ColPlan {
    int off_lo, off_hi;   // byte offsets of the aligned source words inside a grid row
    uint32_t ormask;      // 0xFF in every byte whose column is out of the grid
    int shift;            // 8 * m
};

__device__ __forceinline__ ColPlan make_plan(int G, int j0) {
    ColPlan p;
    const int a = j0 & ~3;                 // aligned start (floor to a multiple of 4, also for negatives)
    const int m = j0 - a;
    p.shift = 8 * m;
    const int last = G - 4;
    p.off_lo = min(max(a, 0), last);
    p.off_hi = min(max(a + 4, 0), last);
    uint32_t om = 0;
#pragma unroll
    for (int c = 0; c < 4; ++c)
        if (static_cast<unsigned>(j0 + c) >= static_cast<unsigned>(G)) om |= 0xFFu << (8 * c);
    p.ormask = om;
    return p;
}

// The per-env scalar step (SPEC.md §7).  WARP = false: one thread per env.  WARP = true: a whole warp runs it
// redundantly for one env (lanes 0..20 take one footprint cell each, lane 0 does the writes); every lane
// terminal observation order of an env that finished (SPEC.md §7): the crop centres of [previous frame, terminal frame] on the
// finished episode's plane; terminal_obs_kernel turns it into info["terminal_local_map"] before the slot is regenerated
__device__ __forceinline__ void store_term_order(const StepArgs &a, int e, int ci, int cj, int pi, int pj, uint32_t episode) {
    int32_t *o = a.term_order + static_cast<size_t>(e) * 8;
    *reinterpret_cast<int4 *>(o) = make_int4(ci, cj, pi, pj);
    o[4] = static_cast<int32_t>((episode % a.S) * static_cast<uint32_t>(a.N) + static_cast<uint32_t>(e));
}

// returns the same crop order.  o0 = (ci, cj, pi, pj), o1 = (plane, flags: bit0 active, bit1 two frames).
template <bool WARP>
__device__ __forceinline__ void dynamics_env(const StepArgs &a, int e, int lane, uint4 &o0, uint2 &o1) {
    const bool lead = !WARP || lane == 0;
    const int G = a.G;
    const size_t cells = static_cast<size_t>(G) * G;
    uint32_t *st = a.state + static_cast<size_t>(e) * ST_WORDS;
    const uint4 s0 = *reinterpret_cast<const uint4 *>(st);        // x y yaw gx
    const uint4 s1 = *reinterpret_cast<const uint4 *>(st + 4);    // gy d_first return steps
    float x = __uint_as_float(s0.x), y = __uint_as_float(s0.y), yaw = __uint_as_float(s0.z);
    const float gx = __uint_as_float(s0.w), gy = __uint_as_float(s1.x);
    const float d_first = __uint_as_float(s1.y);
    float ep_return = __uint_as_float(s1.z);
    int steps = static_cast<int>(s1.w);
    uint32_t episode = st[ST_EPISODE];
    bool begin = false, active = true;
    int ci = 0, cj = 0, pi = 0, pj = 0;

    if (a.mode == 0) {
        long long act = a.actions[e];
        if (act < 0 || act >= 28) {
            act = 3;
            if (lead) atomicOr(a.error_word, 1u);
        }
        float v, w, s, c;
        action_lookup(static_cast<int>(act), v, w);
        sincos_spec(yaw, s, c);
        const float nx = fadd(x, fmul(fmul(v, c), a.dt));
        const float ny = fadd(y, fmul(fmul(v, s), a.dt));
        const float nyaw = pi_to_pi(fadd(yaw, fmul(w, a.dt)));
        ci = robot_cell(nx);
        cj = robot_cell(ny);
        const uint8_t *img = a.flow + (static_cast<size_t>(episode % a.S) * a.N + e) * cells;
        bool col;
        if (WARP) {
            bool hit = false;
            if (lane < 21) {
                const int i = ci + FOOT_DI[lane], j = cj + FOOT_DJ[lane];
                hit = static_cast<unsigned>(i) >= static_cast<unsigned>(G) || static_cast<unsigned>(j) >= static_cast<unsigned>(G);
                if (!hit) hit = __ldg(img + static_cast<size_t>(i) * G + j) == 255;
            }
            col = __ballot_sync(FULL, hit) != 0;
        } else if (ci >= 2 && cj >= 2 && ci < G - 2 && cj < G - 2) {
            // all 21 loads are issued before the first use (one memory round trip, not 21)
            const uint8_t *centre = img + static_cast<size_t>(ci) * G + cj;
            uint32_t cell[21];
#pragma unroll
            for (int k = 0; k < 21; ++k) cell[k] = __ldg(centre + (FOOT_DI[k] * G + FOOT_DJ[k]));
            uint32_t hit = 0;
#pragma unroll
            for (int k = 0; k < 21; ++k) hit |= (cell[k] + 1u) >> 8;   // 255 -> 1, anything else -> 0
            col = hit != 0;
        } else {
            col = true;   // some footprint cell is out of the grid
        }
        const float dx = fsub(gx, nx), dy = fsub(gy, ny);
        const float d = dist_spec(dx, dy);
        const float bearing = pi_to_pi(fsub(atan2_spec(dy, dx), nyaw));
        const float vl = dist_spec(fsub(nx, x), fsub(ny, y));
        const float va = pi_to_pi(fsub(nyaw, yaw));
        const bool goal = d < 0.5f;
        const float r = fadd(fadd(goal ? 1.0f : fmul(0.05f, fsub(d_first, d)), col ? -1.0f : 0.0f), -0.05f);
        steps += 1;
        const bool trunc = steps == a.max_steps;
        const bool done = col || goal || trunc;
        ep_return = fadd(ep_return, r);
        if (lead) {
            a.reward[e] = r;
            a.done[e] = done ? 1 : 0;
            a.flags[e] = static_cast<uint8_t>((col ? 1 : 0) | (goal ? 2 : 0) | (trunc ? 4 : 0));
            *reinterpret_cast<float2 *>(a.term_rel_goal + 2 * e) = make_float2(d, bearing);
            *reinterpret_cast<float2 *>(a.term_velocity + 2 * e) = make_float2(vl, va);
        }
        if (done) {
            if (lead) {
                a.fin_return[e] = ep_return; a.fin_length[e] = steps;
                if (a.term_order) store_term_order(a, e, ci, cj, robot_cell(x), robot_cell(y), episode);
            }
            begin = true;
        } else {
            pi = robot_cell(x);
            pj = robot_cell(y);
            if (lead) {
                st[ST_X] = __float_as_uint(nx); st[ST_Y] = __float_as_uint(ny); st[ST_YAW] = __float_as_uint(nyaw);
                st[ST_RETURN] = __float_as_uint(ep_return);
                st[ST_STEPS] = static_cast<uint32_t>(steps);
                *reinterpret_cast<float2 *>(a.rel_goal + 2 * e) = make_float2(d, bearing);
                *reinterpret_cast<float2 *>(a.velocity + 2 * e) = make_float2(vl, va);
            }
        }
    } else if (a.mode == 1) {
        begin = a.mask[e] != 0;
        active = begin;
    } else {
        begin = true;
    }

    if (begin) {
        if (a.mode != 2) {
            episode += 1;
            if (lead) {
                // the slot of the finished episode is refilled with episode + S - 1
                const uint32_t idx = atomicAdd(a.regen_count, 1u);
                a.regen_env[idx] = static_cast<uint32_t>(e);
                a.regen_episode[idx] = episode + static_cast<uint32_t>(a.S) - 1u;
            }
        }
        const uint32_t *rec = a.scen + (static_cast<size_t>(episode % a.S) * a.N + e) * SC_WORDS;
        const uint4 r0 = *reinterpret_cast<const uint4 *>(rec);
        x = __uint_as_float(r0.x); y = __uint_as_float(r0.y); yaw = __uint_as_float(r0.z);
        const float ngx = __uint_as_float(r0.w), ngy = __uint_as_float(rec[SC_GY]);
        const float dx = fsub(ngx, x), dy = fsub(ngy, y);
        const float d = dist_spec(dx, dy);
        const float bearing = pi_to_pi(fsub(atan2_spec(dy, dx), yaw));
        ci = pi = robot_cell(x);
        cj = pj = robot_cell(y);
        if (lead) {
            *reinterpret_cast<uint4 *>(st) = make_uint4(__float_as_uint(x), __float_as_uint(y), __float_as_uint(yaw), __float_as_uint(ngx));
            *reinterpret_cast<uint4 *>(st + 4) = make_uint4(__float_as_uint(ngy), __float_as_uint(d), __float_as_uint(0.0f), 0u);
            st[ST_EPISODE] = episode;
            *reinterpret_cast<float2 *>(a.rel_goal + 2 * e) = make_float2(d, bearing);
            *reinterpret_cast<float2 *>(a.velocity + 2 * e) = make_float2(0.0f, 0.0f);
            if (a.mode == 2) { a.reward[e] = 0.0f; a.done[e] = 0; a.flags[e] = 0; }
        }
    }
    o0 = make_uint4(static_cast<uint32_t>(ci), static_cast<uint32_t>(cj), static_cast<uint32_t>(pi), static_cast<uint32_t>(pj));
    o1 = make_uint2((episode % a.S) * static_cast<uint32_t>(a.N) + static_cast<uint32_t>(e),
                    (active ? 1u : 0u) | ((begin || a.write_older) ? 2u : 0u));
}

__global__ void __launch_bounds__(128) dynamics_kernel(StepArgs a) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= a.N) return;
    uint4 o0;
    uint2 o1;
    dynamics_env<false>(a, e, 0, o0, o1);
    uint32_t *ob = a.obs_order + static_cast<size_t>(e) * OB_WORDS;
    *reinterpret_cast<uint4 *>(ob) = o0;
    *reinterpret_cast<uint2 *>(ob + 4) = o1;
}

constexpr int ROWS_IN_FLIGHT = 8;

__device__ __forceinline__ int ceil_div_pos(int a, int b) { return a <= 0 ? 0 : (a + b - 1) / b; }

// One frame for one lane: rows row0, row0+stride, ... < W of output word column `wl`.
// The lane's rows are split into [out-of-grid | in-grid | out-of-grid] so that the in-grid loop carries no
// bounds checks; it indexes the image as 32-bit words (one IMAD.WIDE per access) and keeps
// ROWS_IN_FLIGHT rows of loads in flight before the first store.
template <bool TWO>
__device__ __forceinline__ void crop_rows(const uint8_t *__restrict__ img, int G, int W, int wpr, int i0, int row0,
                                          int stride, int wl, const ColPlan &p, uint32_t *__restrict__ dst,
                                          uint32_t *__restrict__ dst2) {
    const int nk = ceil_div_pos(W - row0, stride);               // rows owned by this lane
    const int k_lo = min(ceil_div_pos(-i0 - row0, stride), nk);  // first k with i0 + row >= 0
    const int k_hi = min(ceil_div_pos(G - i0 - row0, stride), nk);   // first k with i0 + row >= G
    const int dstep = stride * wpr;
    int d = row0 * wpr + wl;
    for (int k = 0; k < k_lo; ++k, d += dstep) {
        __stcs(dst + d, 0xFFFFFFFFu);
        if (TWO) __stcs(dst2 + d, 0xFFFFFFFFu);
    }
    const uint32_t *img32 = reinterpret_cast<const uint32_t *>(img);
    const int sstep = (stride * G) >> 2;
    int s_lo = ((i0 + row0 + k_lo * stride) * G + p.off_lo) >> 2;
    int s_hi = ((i0 + row0 + k_lo * stride) * G + p.off_hi) >> 2;
    int k = k_lo;
    for (; k + ROWS_IN_FLIGHT <= k_hi; k += ROWS_IN_FLIGHT) {
        uint32_t lo[ROWS_IN_FLIGHT], hi[ROWS_IN_FLIGHT];
#pragma unroll
        for (int u = 0; u < ROWS_IN_FLIGHT; ++u) {
            lo[u] = __ldg(img32 + s_lo + u * sstep);
            hi[u] = __ldg(img32 + s_hi + u * sstep);
        }
#pragma unroll
        for (int u = 0; u < ROWS_IN_FLIGHT; ++u) {
            const uint32_t v = __funnelshift_r(lo[u], hi[u], p.shift) | p.ormask;
            __stcs(dst + d + u * dstep, v);
            if (TWO) __stcs(dst2 + d + u * dstep, v);
        }
        s_lo += ROWS_IN_FLIGHT * sstep; s_hi += ROWS_IN_FLIGHT * sstep; d += ROWS_IN_FLIGHT * dstep;
    }
    {   // tail: fewer than ROWS_IN_FLIGHT in-grid rows left, still issued as one batch of loads
        uint32_t lo[ROWS_IN_FLIGHT], hi[ROWS_IN_FLIGHT];
        const int rem = k_hi - k;
#pragma unroll
        for (int u = 0; u < ROWS_IN_FLIGHT - 1; ++u)
            if (u < rem) {
                lo[u] = __ldg(img32 + s_lo + u * sstep);
                hi[u] = __ldg(img32 + s_hi + u * sstep);
            }
#pragma unroll
        for (int u = 0; u < ROWS_IN_FLIGHT - 1; ++u)
            if (u < rem) {
                const uint32_t v = __funnelshift_r(lo[u], hi[u], p.shift) | p.ormask;
                __stcs(dst + d + u * dstep, v);
                if (TWO) __stcs(dst2 + d + u * dstep, v);
            }
        d += rem * dstep;
    }
    for (k = k_hi; k < nk; ++k, d += dstep) {
        __stcs(dst + d, 0xFFFFFFFFu);
        if (TWO) __stcs(dst2 + d, 0xFFFFFFFFu);
    }
}

__global__ void __launch_bounds__(128) observe_kernel(StepArgs a) {
    const int e = blockIdx.x;
    const int tid = threadIdx.x;
    const int G = a.G, W = a.W;
    const int wpr = W >> 2;                                  // output words per row
    const int lane = tid & 31, warp = tid >> 5;
    // everything above is independent of dynamics_kernel; its results are consumed below this point
    asm volatile("griddepcontrol.wait;" ::: "memory");
    const uint32_t *ob = a.obs_order + static_cast<size_t>(e) * OB_WORDS;
    const uint4 o0 = *reinterpret_cast<const uint4 *>(ob);
    const uint2 o1 = *reinterpret_cast<const uint2 *>(ob + 4);
    if (!(o1.y & 1u)) return;
    const bool two = (o1.y & 2u) != 0;
    const int ci = static_cast<int>(o0.x), cj = static_cast<int>(o0.y), pi = static_cast<int>(o0.z), pj = static_cast<int>(o0.w);
    const uint8_t *img = a.flow + static_cast<size_t>(o1.x) * G * G;
    uint32_t *f_new = reinterpret_cast<uint32_t *>(a.frames + (static_cast<size_t>(e) * a.K + a.slot_new) * W * W);
    uint32_t *f_old = f_new - (W * W >> 2);
    const bool same = pi == ci && pj == cj;
    const int i0 = ci - (W >> 1), j0 = cj - (W >> 1);
    const int p0 = pi - (W >> 1), q0 = pj - (W >> 1);
    if (wpr <= 32) {
        const int rpi = 32 / wpr;                            // rows per warp iteration
        const int sub = lane / wpr, wl = lane - sub * wpr;
        if (sub >= rpi) return;
        const ColPlan pn = make_plan(G, j0 + 4 * wl);
        if (two && same) {
            crop_rows<true>(img, G, W, wpr, i0, warp * rpi + sub, 4 * rpi, wl, pn, f_new, f_old);
        } else {
            crop_rows<false>(img, G, W, wpr, i0, warp * rpi + sub, 4 * rpi, wl, pn, f_new, nullptr);
            if (two) {
                const ColPlan po = make_plan(G, q0 + 4 * wl);
                crop_rows<false>(img, G, W, wpr, p0, warp * rpi + sub, 4 * rpi, wl, po, f_old, nullptr);
            }
        }
    } else {
        for (int wl = lane; wl < wpr; wl += 32) {
            const ColPlan pn = make_plan(G, j0 + 4 * wl);
            crop_rows<false>(img, G, W, wpr, i0, warp, 4, wl, pn, f_new, nullptr);
            if (two) {
                const ColPlan po = make_plan(G, q0 + 4 * wl);
                crop_rows<false>(img, G, W, wpr, p0, warp, 4, wl, po, f_old, nullptr);
            }
        }
    }
}

// ---- TMA variant of observe_kernel (G % 16 == 0) -------------------------------------------------------
// One `cp.async.bulk.tensor.3d` per frame pulls the W-row window of the flow image into shared memory: the
// TMA engine does the address generation and keeps ~13 KB per CTA in flight without a single register.  TMA
// needs a 16-byte aligned inner coordinate, so the box starts at floor16(j0) and is ceil16(W+15) bytes wide;
// the remaining 0..15 byte misalignment is removed while draining the tile (two LDS + funnel shift).
// Out-of-grid bytes arrive as zeros and are turned into 255 with a per-lane column mask and a per-row mask.
// The CTA streams the dense W x W frame to the ring with coalesced 32-bit stores.
__device__ __forceinline__ void tma_window(uint32_t dst, const CUtensorMap *tm, int c0, int c1, int c2, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                 ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tm)), "r"(c0), "r"(c1), "r"(c2), "r"(bar) : "memory");
}

// One frame for one lane out of the staged tile: rows row0, row0+stride, ... < W of output word column `wl`.  The rows
// are walked as [out-of-grid | in-grid | out-of-grid] so that the in-grid loop is two LDS, one funnel shift, one OR
// (out-of-grid columns -> 255) and the store(s) per row, five rows in flight.
template <bool TWO>
__device__ __forceinline__ void drain_tile(const uint32_t *__restrict__ tile32, int tile_wpr, int G, int W, int wpr,
                                           int i0, int mis, int row0, int stride, int wl, uint32_t ormask,
                                           uint32_t *__restrict__ dst, uint32_t *__restrict__ dst2, int row_end = -1) {
    constexpr int U = 5;
    const int RE = row_end < 0 ? W : row_end;                 // this call covers rows [row0, RE) of the frame
    const int sstep = stride * tile_wpr, dstep = stride * wpr;
    const int shift = 8 * (mis & 3);
    const int r_in = max(-i0, 0), r_out = min(G - i0, RE);     // in-grid rows are [r_in, r_out)
    int row = row0, s = row0 * tile_wpr + (mis >> 2) + wl, d = row0 * wpr + wl;
    for (; row < min(r_in, RE); row += stride, s += sstep, d += dstep) {
        __stcs(dst + d, 0xFFFFFFFFu);
        if (TWO) __stcs(dst2 + d, 0xFFFFFFFFu);
    }
    for (; row + (U - 1) * stride < r_out; row += U * stride, s += U * sstep, d += U * dstep) {
        uint32_t lo[U], hi[U];
#pragma unroll
        for (int u = 0; u < U; ++u) { lo[u] = tile32[s + u * sstep]; hi[u] = tile32[s + u * sstep + 1]; }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const uint32_t v = __funnelshift_r(lo[u], hi[u], shift) | ormask;
            __stcs(dst + d + u * dstep, v);
            if (TWO) __stcs(dst2 + d + u * dstep, v);
        }
    }
    for (; row < r_out; row += stride, s += sstep, d += dstep) {
        const uint32_t v = __funnelshift_r(tile32[s], tile32[s + 1], shift) | ormask;
        __stcs(dst + d, v);
        if (TWO) __stcs(dst2 + d, v);
    }
    for (; row < RE; row += stride, d += dstep) {
        __stcs(dst + d, 0xFFFFFFFFu);
        if (TWO) __stcs(dst2 + d, 0xFFFFFFFFu);
    }
}

// FUSED = true: warp 0 runs the scalar step of this env itself (dynamics_env<true>) before issuing the TMA, so a
// tick is ONE kernel; the other warps sleep on the barrier meanwhile and other CTAs of the SM cover the latency.
// FUSED = false: the crop order comes from dynamics_kernel (programmatic dependent launch).
template <bool FUSED>
__global__ void __launch_bounds__(128) observe_tma_kernel(const __grid_constant__ CUtensorMap tmap, StepArgs a) {
    extern __shared__ __align__(128) uint8_t tile[];
    __shared__ __align__(8) uint64_t mbar;
    const int e = blockIdx.x;
    const int tid = threadIdx.x;
    const int G = a.G, W = a.W;
    const int wpr = W >> 2;
    const int tile_w = (W + 15 + 15) & ~15;                   // TMA box inner extent: W + worst misalignment, multiple of 16
    const int lane = tid & 31, warp = tid >> 5;
    const uint32_t bar = static_cast<uint32_t>(__cvta_generic_to_shared(&mbar));
    const uint32_t tile_s = static_cast<uint32_t>(__cvta_generic_to_shared(tile));
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    uint4 o0;
    uint2 o1;
    if (FUSED) {
        __shared__ uint4 s_o0;
        __shared__ uint2 s_o1;
        if (warp == 0) {
            dynamics_env<true>(a, e, lane, o0, o1);
            if (lane == 0) { s_o0 = o0; s_o1 = o1; }
        }
        __syncthreads();
        o0 = s_o0; o1 = s_o1;
    } else {
        __syncthreads();
        asm volatile("griddepcontrol.wait;" ::: "memory");
        const uint32_t *ob = a.obs_order + static_cast<size_t>(e) * OB_WORDS;
        o0 = *reinterpret_cast<const uint4 *>(ob);
        o1 = *reinterpret_cast<const uint2 *>(ob + 4);
    }
    if (!(o1.y & 1u)) return;
    const bool two = (o1.y & 2u) != 0;
    const int ci = static_cast<int>(o0.x), cj = static_cast<int>(o0.y), pi = static_cast<int>(o0.z), pj = static_cast<int>(o0.w);
    const bool same = pi == ci && pj == cj;
    const int i0 = ci - (W >> 1), j0 = cj - (W >> 1);
    const uint32_t bytes = static_cast<uint32_t>(tile_w * W);
    if (tid == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
        tma_window(tile_s, &tmap, j0 & ~15, i0, static_cast<int>(o1.x), bar);
    }
    uint32_t *f_new = reinterpret_cast<uint32_t *>(a.frames + (static_cast<size_t>(e) * a.K + a.slot_new) * W * W);
    uint32_t *f_old = f_new - (W * W >> 2);
    // lane -> (row slot, word) mapping and the column mask (out-of-grid columns -> 255)
    int rpi = 1, sub = 0, wl = lane;
    if (wpr <= 32) { rpi = 32 / wpr; sub = lane / wpr; wl = lane - sub * wpr; }
    auto colmask = [&](int jbase, int w) {
        uint32_t om = 0;
#pragma unroll
        for (int c = 0; c < 4; ++c)
            if (static_cast<unsigned>(jbase + 4 * w + c) >= static_cast<unsigned>(G)) om |= 0xFFu << (8 * c);
        return om;
    };
    const uint32_t *tile32 = reinterpret_cast<const uint32_t *>(tile);
    const int tile_wpr = tile_w >> 2;
    // wait for the window (phase 0)
    {
        uint32_t done = 0;
        while (!done)
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(done) : "r"(bar), "r"(0u) : "memory");
    }
    if (wpr <= 32) {
        if (sub < rpi) {
            const uint32_t om = colmask(j0, wl);
            if (two && same) drain_tile<true>(tile32, tile_wpr, G, W, wpr, i0, j0 & 15, warp * rpi + sub, 4 * rpi, wl, om, f_new, f_old);
            else drain_tile<false>(tile32, tile_wpr, G, W, wpr, i0, j0 & 15, warp * rpi + sub, 4 * rpi, wl, om, f_new, nullptr);
        }
    } else {
        for (int w2 = lane; w2 < wpr; w2 += 32)
            drain_tile<false>(tile32, tile_wpr, G, W, wpr, i0, j0 & 15, warp, 4, w2, colmask(j0, w2), f_new, nullptr);
        if (two && same)
            for (int w2 = lane; w2 < wpr; w2 += 32)
                drain_tile<false>(tile32, tile_wpr, G, W, wpr, i0, j0 & 15, warp, 4, w2, colmask(j0, w2), f_old, nullptr);
    }
    if (two && !same) {
        // older frame of a non-reset env on a ring wrap: the window at the previous pose
        const int p0 = pi - (W >> 1), q0 = pj - (W >> 1);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic reads of `tile` before the async overwrite
        __syncthreads();
        if (tid == 0) {
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
            tma_window(tile_s, &tmap, q0 & ~15, p0, static_cast<int>(o1.x), bar);
        }
        uint32_t done = 0;
        while (!done)
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(done) : "r"(bar), "r"(1u) : "memory");
        if (wpr <= 32) {
            if (sub < rpi)
                drain_tile<false>(tile32, tile_wpr, G, W, wpr, p0, q0 & 15, warp * rpi + sub, 4 * rpi, wl, colmask(q0, wl), f_old, nullptr);
        } else {
            for (int w2 = lane; w2 < wpr; w2 += 32)
                drain_tile<false>(tile32, tile_wpr, G, W, wpr, p0, q0 & 15, warp, 4, w2, colmask(q0, w2), f_old, nullptr);
        }
    }
}

// ---- tick_tma_kernel: the whole env step in ONE kernel, latency-optimised --------------------------------
// One CTA per env.  Warp 0: state + action -> kinematics -> robot cell, then IMMEDIATELY issues the TMA for the
// window around the new cell (speculating that the episode continues) and overlaps the relative-goal / velocity
// math with the copy.  When the tile has landed the 21 footprint cells are read from SHARED memory (they sit
// at the centre of the window), so the collision test costs no extra global round trip.  Only if the episode
// ended (rare) a second TMA fetches the first window of the next scenario.  All four warps then stream the
// frame to the ring.  Critical path per CTA: one state load, one TMA, one drain.
struct TickShared {
    int go, two, ci, cj, pi, pj, parity;
    unsigned int plane;
};

__device__ __forceinline__ void mbar_wait_parity(uint32_t bar, uint32_t parity) {
    uint32_t done = 0;
    while (!done)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(bar), "r"(parity) : "memory");
}

// WT != 0: the window side is a compile-time constant (the reference's 100): row pitches and unrolled store offsets become
// immediates, which takes ~40 % of the instructions out of the drain loop.  WT = 0: generic window (a.W).
// PA: the actions travel IN the launch, as a by-value kernel parameter of ACT_PARAM_MAX bytes (one byte per env, 255 = invalid),
// and are read from the constant bank: a host-buffer step (ffmp_step_host) then needs neither a copy engine in front of the
// kernel nor a PCIe read inside every CTA's dependent chain.
template <bool PA> struct ActionBlock { uint8_t a[PA ? ACT_PARAM_MAX : 4]; };

template <bool TRACE, int WT, bool PA = false>
__global__ void __launch_bounds__(128) tick_tma_kernel(const __grid_constant__ CUtensorMap tmap, StepArgs a,
                                                       const __grid_constant__ ActionBlock<PA> ab) {
    extern __shared__ __align__(128) uint8_t tile[];
    __shared__ __align__(8) uint64_t mbar;
    __shared__ TickShared sh;
    const int e = blockIdx.x;
    const int tid = threadIdx.x;
    const int G = a.G, W = WT ? WT : a.W;
    const int wpr = W >> 2;
    const int tile_w = (W + 15 + 15) & ~15;
    const int tile_wpr = tile_w >> 2;
    const int lane = tid & 31, warp = tid >> 5;
    const uint32_t bar = static_cast<uint32_t>(__cvta_generic_to_shared(&mbar));
    const uint32_t tile_s = static_cast<uint32_t>(__cvta_generic_to_shared(tile));
    const uint32_t bytes = static_cast<uint32_t>(tile_w * W);
    const int half = W >> 1;
    unsigned long long *tr = TRACE ? a.trace + static_cast<size_t>(e) * 8 : nullptr;
    auto stamp = [&](int slot) {
        if (TRACE && tid == 0) {
            if (slot == 0 || slot == 7) {
                unsigned long long g;
                asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g));
                tr[slot] = g;
            } else {
                tr[slot] = static_cast<unsigned long long>(clock64());
            }
        }
    };
    stamp(0);
    stamp(1);
    // a programmatic dependent of this launch (host_export_kernel, host_io.cu) may become resident once every CTA of the
    // tick has started; it still waits for the whole grid (griddepcontrol.wait) before it reads the results
    asm volatile("griddepcontrol.launch_dependents;");

    if (warp == 0) {
        if (lane == 0) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar) : "memory");
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncwarp();
        // when this tick was launched as a programmatic dependent of the previous one its CTAs are resident early and
        // wait here for the previous grid (returns at once otherwise)
        asm volatile("griddepcontrol.wait;" ::: "memory");
        uint32_t *st = a.state + static_cast<size_t>(e) * ST_WORDS;
        const uint4 s0 = *reinterpret_cast<const uint4 *>(st);        // x y yaw gx
        const uint4 s1 = *reinterpret_cast<const uint4 *>(st + 4);    // gy d_first return steps
        float x = __uint_as_float(s0.x), y = __uint_as_float(s0.y), yaw = __uint_as_float(s0.z);
        const float gx = __uint_as_float(s0.w), gy = __uint_as_float(s1.x);
        const float d_first = __uint_as_float(s1.y);
        float ep_return = __uint_as_float(s1.z);
        int steps = static_cast<int>(s1.w);
        uint32_t episode = st[ST_EPISODE];
        bool begin = false, active = true;
        int ci = 0, cj = 0, pi = 0, pj = 0;
        int parity = 0;

        if (a.mode == 0) {
            long long act = PA ? static_cast<long long>(ab.a[e]) : a.actions[e];
            if (act < 0 || act >= 28) {
                act = 3;
                if (lane == 0) atomicOr(a.error_word, 1u);
            }
            float v, w, s, c;
            action_lookup(static_cast<int>(act), v, w);
            sincos_spec(yaw, s, c);
            const float nx = fadd(x, fmul(fmul(v, c), a.dt));
            const float ny = fadd(y, fmul(fmul(v, s), a.dt));
            const float nyaw = pi_to_pi(fadd(yaw, fmul(w, a.dt)));
            ci = robot_cell(nx);
            cj = robot_cell(ny);
            const int i0 = ci - half, j0 = cj - half;
            stamp(2);
            if (lane == 0) {   // speculative window fetch: the episode usually continues
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
                tma_window(tile_s, &tmap, j0 & ~15, i0, static_cast<int>((episode % a.S) * a.N + e), bar);
            }
            // math that does not depend on the map overlaps the copy
            const float dx = fsub(gx, nx), dy = fsub(gy, ny);
            const float d = dist_spec(dx, dy);
            const float bearing = pi_to_pi(fsub(atan2_spec(dy, dx), nyaw));
            const float vl = dist_spec(fsub(nx, x), fsub(ny, y));
            const float va = pi_to_pi(fsub(nyaw, yaw));
            const bool goal = d < 0.5f;
            stamp(3);
            mbar_wait_parity(bar, 0);
            stamp(4);
            bool hit = false;
            if (lane < 21) {
                // footprint offset of this lane from the lane index (a per-lane __constant__ lookup would be serialised)
                int di, dj;
                if (lane < 3) { di = -2; dj = lane - 1; }
                else if (lane < 18) { di = (lane - 3) / 5 - 1; dj = (lane - 3) % 5 - 2; }
                else { di = 2; dj = lane - 19; }
                const int i = ci + di, j = cj + dj;
                hit = static_cast<unsigned>(i) >= static_cast<unsigned>(G) || static_cast<unsigned>(j) >= static_cast<unsigned>(G);
                if (!hit) hit = tile[(half + di) * tile_w + (j0 & 15) + half + dj] == 255;
            }
            const bool col = __ballot_sync(FULL, hit) != 0;
            const float r = fadd(fadd(goal ? 1.0f : fmul(0.05f, fsub(d_first, d)), col ? -1.0f : 0.0f), -0.05f);
            steps += 1;
            const bool trunc = steps == a.max_steps;
            const bool done = col || goal || trunc;
            ep_return = fadd(ep_return, r);
            if (lane == 0) {
                a.reward[e] = r;
                a.done[e] = done ? 1 : 0;
                a.flags[e] = static_cast<uint8_t>((col ? 1 : 0) | (goal ? 2 : 0) | (trunc ? 4 : 0));
                *reinterpret_cast<float2 *>(a.term_rel_goal + 2 * e) = make_float2(d, bearing);
                *reinterpret_cast<float2 *>(a.term_velocity + 2 * e) = make_float2(vl, va);
            }
            if (done) {
                if (lane == 0) {
                    a.fin_return[e] = ep_return; a.fin_length[e] = steps;
                    if (a.term_order) store_term_order(a, e, ci, cj, robot_cell(x), robot_cell(y), episode);
                }
                begin = true;
                parity = 1;   // the window of the next scenario arrives in the barrier's second phase
            } else {
                pi = robot_cell(x);
                pj = robot_cell(y);
                if (lane == 0) {
                    st[ST_X] = __float_as_uint(nx); st[ST_Y] = __float_as_uint(ny); st[ST_YAW] = __float_as_uint(nyaw);
                    st[ST_RETURN] = __float_as_uint(ep_return);
                    st[ST_STEPS] = static_cast<uint32_t>(steps);
                    *reinterpret_cast<float2 *>(a.rel_goal + 2 * e) = make_float2(d, bearing);
                    *reinterpret_cast<float2 *>(a.velocity + 2 * e) = make_float2(vl, va);
                }
            }
        } else if (a.mode == 1) {
            begin = a.mask[e] != 0;
            active = begin;
        } else {
            begin = true;
        }

        if (begin) {
            if (a.mode != 2) {
                episode += 1;
                if (lane == 0) {
                    const uint32_t idx = atomicAdd(a.regen_count, 1u);
                    a.regen_env[idx] = static_cast<uint32_t>(e);
                    a.regen_episode[idx] = episode + static_cast<uint32_t>(a.S) - 1u;
                }
            }
            const uint32_t *rec = a.scen + (static_cast<size_t>(episode % a.S) * a.N + e) * SC_WORDS;
            const uint4 r0 = *reinterpret_cast<const uint4 *>(rec);
            x = __uint_as_float(r0.x); y = __uint_as_float(r0.y); yaw = __uint_as_float(r0.z);
            const float ngx = __uint_as_float(r0.w), ngy = __uint_as_float(rec[SC_GY]);
            ci = pi = robot_cell(x);
            cj = pj = robot_cell(y);
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // footprint reads of `tile` precede the overwrite
            __syncwarp();
            if (lane == 0) {
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
                tma_window(tile_s, &tmap, (cj - half) & ~15, ci - half, static_cast<int>((episode % a.S) * a.N + e), bar);
            }
            const float dx = fsub(ngx, x), dy = fsub(ngy, y);
            const float d = dist_spec(dx, dy);
            const float bearing = pi_to_pi(fsub(atan2_spec(dy, dx), yaw));
            if (lane == 0) {
                *reinterpret_cast<uint4 *>(st) = make_uint4(__float_as_uint(x), __float_as_uint(y), __float_as_uint(yaw), __float_as_uint(ngx));
                *reinterpret_cast<uint4 *>(st + 4) = make_uint4(__float_as_uint(ngy), __float_as_uint(d), __float_as_uint(0.0f), 0u);
                st[ST_EPISODE] = episode;
                *reinterpret_cast<float2 *>(a.rel_goal + 2 * e) = make_float2(d, bearing);
                *reinterpret_cast<float2 *>(a.velocity + 2 * e) = make_float2(0.0f, 0.0f);
                if (a.mode == 2) { a.reward[e] = 0.0f; a.done[e] = 0; a.flags[e] = 0; }
            }
        }
        if (lane == 0) {
            sh.go = active ? 1 : 0;
            sh.two = (begin || a.write_older) ? 1 : 0;
            sh.ci = ci; sh.cj = cj; sh.pi = pi; sh.pj = pj;
            sh.parity = parity;
            sh.plane = (episode % a.S) * static_cast<uint32_t>(a.N) + static_cast<uint32_t>(e);
        }
    }
    __syncthreads();
    if (!sh.go) return;
    const bool two = sh.two != 0;
    const int ci = sh.ci, cj = sh.cj, pi = sh.pi, pj = sh.pj;
    const bool same = pi == ci && pj == cj;
    const int i0 = ci - half, j0 = cj - half;
    uint32_t parity = static_cast<uint32_t>(sh.parity);
    mbar_wait_parity(bar, parity);
    stamp(5);

    uint32_t *f_new = reinterpret_cast<uint32_t *>(a.frames + (static_cast<size_t>(e) * a.K + a.slot_new) * W * W);
    uint32_t *f_old = f_new - (W * W >> 2);
    int rpi = 1, sub = 0, wl = lane;
    if (wpr <= 32) { rpi = 32 / wpr; sub = lane / wpr; wl = lane - sub * wpr; }
    auto colmask = [&](int jbase, int w) {
        uint32_t om = 0;
#pragma unroll
        for (int c = 0; c < 4; ++c)
            if (static_cast<unsigned>(jbase + 4 * w + c) >= static_cast<unsigned>(G)) om |= 0xFFu << (8 * c);
        return om;
    };
    const uint32_t *tile32 = reinterpret_cast<const uint32_t *>(tile);
    if (wpr <= 32) {
        if (sub < rpi) {
            const uint32_t om = colmask(j0, wl);
            if (two && same) drain_tile<true>(tile32, tile_wpr, G, W, wpr, i0, j0 & 15, warp * rpi + sub, 4 * rpi, wl, om, f_new, f_old);
            else drain_tile<false>(tile32, tile_wpr, G, W, wpr, i0, j0 & 15, warp * rpi + sub, 4 * rpi, wl, om, f_new, nullptr);
        }
    } else {
        for (int w2 = lane; w2 < wpr; w2 += 32)
            drain_tile<false>(tile32, tile_wpr, G, W, wpr, i0, j0 & 15, warp, 4, w2, colmask(j0, w2), f_new, nullptr);
        if (two && same)
            for (int w2 = lane; w2 < wpr; w2 += 32)
                drain_tile<false>(tile32, tile_wpr, G, W, wpr, i0, j0 & 15, warp, 4, w2, colmask(j0, w2), f_old, nullptr);
    }
    if (two && !same) {
        // ring wrap of a continuing env: the older frame is the window at the previous pose
        const int p0 = pi - half, q0 = pj - half;
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();
        if (tid == 0) {
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
            tma_window(tile_s, &tmap, q0 & ~15, p0, static_cast<int>(sh.plane), bar);
        }
        mbar_wait_parity(bar, parity ^ 1u);
        if (wpr <= 32) {
            if (sub < rpi)
                drain_tile<false>(tile32, tile_wpr, G, W, wpr, p0, q0 & 15, warp * rpi + sub, 4 * rpi, wl, colmask(q0, wl), f_old, nullptr);
        } else {
            for (int w2 = lane; w2 < wpr; w2 += 32)
                drain_tile<false>(tile32, tile_wpr, G, W, wpr, p0, q0 & 15, warp, 4, w2, colmask(q0, w2), f_old, nullptr);
        }
    }
    __syncthreads();
    stamp(6);
    stamp(7);
    if (TRACE && tid == 0) {
        unsigned int smid;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        tr[1] = (tr[1] << 8) | smid;     // clock at start (shifted) + SM id in the low byte
    }
}

// Batched FFMP.rewarder / rewarder2 / reward_calculator (ffmp.py:130-188): one warp per item.
__global__ void __launch_bounds__(128) rewarder_kernel(RewarderArgs a) {
    const int item = blockIdx.x * 4 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (item >= a.n) return;
    bool hit = false;
    if (a.local_map) {
        const int W = a.W, c = W >> 1;
        if (lane < 21) {
            const int i = c + FOOT_DI[lane], j = c + FOOT_DJ[lane];
            if (i >= 0 && j >= 0 && i < W && j < W) hit = a.local_map[(static_cast<size_t>(item) * W + i) * W + j] > 0;
        }
    } else if (a.scan) {
        for (int k = lane; k < a.scan_len; k += 32) {
            const double r = a.scan[static_cast<size_t>(item) * a.scan_len + k];
            // `if scan[i]: if scan[i] < 0.13` on Python floats (fp64), None encoded as NaN
            if (r == r && r != 0.0 && r < 0.13) hit = true;
        }
    }
    bool col = __ballot_sync(FULL, hit) != 0;
    if (lane == 0) {
        const float d = a.rel_goal[2 * item];
        if (a.is_first[item]) a.d_first[item] = d;
        const float d_first = a.d_first[item];
        bool goal = d < 0.5f;
        if (a.given_flags) { col = a.given_flags[item] & 1; goal = a.given_flags[item] & 2; }
        a.reward[item] = fadd(fadd(goal ? 1.0f : fmul(0.05f, fsub(d_first, d)), col ? -1.0f : 0.0f), -0.05f);
        a.done[item] = (col || goal) ? 1 : 0;
        a.flags[item] = static_cast<uint8_t>((col ? 1 : 0) | (goal ? 2 : 0));
    }
}

}  // namespace

// Dynamic shared memory above 48 KB needs a per-function opt-in, and the attribute belongs to the device (context) the call
// is made on: one process may drive several GPUs through the C-ABI, so the opt-in is tracked per device ordinal.
static cudaError_t opt_in_shared(size_t smem) {
    static std::mutex mu;
    static size_t configured[64] = {0};
    int dev = 0;
    cudaError_t ce = cudaGetDevice(&dev);
    if (ce != cudaSuccess) return ce;
    std::lock_guard<std::mutex> lock(mu);
    if (dev >= 0 && dev < 64 && smem <= configured[dev]) return cudaSuccess;
    ce = cudaFuncSetAttribute(tick_tma_kernel<false, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (ce == cudaSuccess)
        ce = cudaFuncSetAttribute(tick_tma_kernel<false, 0, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (ce == cudaSuccess)
        ce = cudaFuncSetAttribute(tick_tma_kernel<true, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (ce == cudaSuccess)
        ce = cudaFuncSetAttribute(observe_tma_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (ce == cudaSuccess && dev >= 0 && dev < 64) configured[dev] = smem;
    return ce;
}

cudaError_t launch_step(const StepArgs &a, const CUtensorMap *tmap, cudaStream_t st, cudaEvent_t between, bool fused, bool pdl,
                        const uint8_t *act_bytes) {
    if (a.N <= 0) return cudaSuccess;
    const size_t smem = tmap ? static_cast<size_t>((a.W + 30) & ~15) * a.W + 16 : 0;
    if (tmap) {
        if (smem > 48 * 1024) {
            const cudaError_t ce = opt_in_shared(smem);
            if (ce != cudaSuccess) return ce;
        }
        if (fused && !between) {
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3(static_cast<unsigned>(a.N));
            cfg.blockDim = dim3(128);
            cfg.stream = st;
            cfg.dynamicSmemBytes = smem;
            cudaLaunchAttribute attr[1];
            attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
            attr[0].val.programmaticStreamSerializationAllowed = 1;
            cfg.attrs = attr;
            cfg.numAttrs = pdl ? 1 : 0;
            if (act_bytes && !a.trace && a.N <= ACT_PARAM_MAX) {
                // actions by value (host-buffer steps of up to ACT_PARAM_MAX envs): the bytes are copied into the launch here
                ActionBlock<true> ab;
                std::memcpy(ab.a, act_bytes, static_cast<size_t>(a.N));
                if (a.W == 100) return cudaLaunchKernelEx(&cfg, tick_tma_kernel<false, 100, true>, *tmap, a, ab);
                if (a.W == 64) return cudaLaunchKernelEx(&cfg, tick_tma_kernel<false, 64, true>, *tmap, a, ab);
                return cudaLaunchKernelEx(&cfg, tick_tma_kernel<false, 0, true>, *tmap, a, ab);
            }
            const ActionBlock<false> none{};
            if (a.trace) return cudaLaunchKernelEx(&cfg, tick_tma_kernel<true, 0>, *tmap, a, none);
            if (a.W == 100) return cudaLaunchKernelEx(&cfg, tick_tma_kernel<false, 100>, *tmap, a, none);   // the reference's window
            if (a.W == 64) return cudaLaunchKernelEx(&cfg, tick_tma_kernel<false, 64>, *tmap, a, none);     // BASELINE config 2
            return cudaLaunchKernelEx(&cfg, tick_tma_kernel<false, 0>, *tmap, a, none);   // the whole tick in one kernel
        }
    }
    dynamics_kernel<<<(a.N + 127) / 128, 128, 0, st>>>(a);
    cudaError_t ce = cudaGetLastError();
    if (ce != cudaSuccess) return ce;
    if (between) {   // profiling mode: a timing event between the two kernels
        ce = cudaEventRecord(between, st);
        if (ce != cudaSuccess) return ce;
    }
    // programmatic dependent launch: the observe CTAs become resident while dynamics_kernel drains
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(static_cast<unsigned>(a.N));
    cfg.blockDim = dim3(128);
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cfg.dynamicSmemBytes = smem;
    if (tmap) return cudaLaunchKernelEx(&cfg, observe_tma_kernel<false>, *tmap, a);
    return cudaLaunchKernelEx(&cfg, observe_kernel, a);
}

// info["terminal_local_map"] (SPEC.md §7, a15; the reference's last tick of an episode, train.py:611-664): one CTA per env,
// envs that did not finish exit at once (~2.5 % finish per step in the bench workload).  Runs on the step's stream before the
// finished slot's regeneration is released, so the finished episode's flow image is still intact.
__global__ void __launch_bounds__(128) terminal_obs_kernel(StepArgs a) {
    const int e = blockIdx.x;
    if (a.done[e] == 0) return;
    const int G = a.G, W = a.W, half = W >> 1;
    const int32_t *o = a.term_order + static_cast<size_t>(e) * 8;
    const uint8_t *img = a.flow + static_cast<size_t>(static_cast<uint32_t>(o[4])) * G * G;
    uint8_t *dst = a.term_frames + static_cast<size_t>(e) * 2 * W * W;
    for (int f = 0; f < 2; ++f) {
        const int i0 = (f == 0 ? o[2] : o[0]) - half, j0 = (f == 0 ? o[3] : o[1]) - half;
        for (int k = threadIdx.x; k < W * W; k += blockDim.x) {
            const int r = k / W, c = k - r * W;
            const int i = i0 + r, j = j0 + c;
            const bool in = static_cast<unsigned>(i) < static_cast<unsigned>(G) && static_cast<unsigned>(j) < static_cast<unsigned>(G);
            dst[static_cast<size_t>(f) * W * W + k] = in ? __ldg(img + static_cast<size_t>(i) * G + j) : static_cast<uint8_t>(255);
        }
    }
}

cudaError_t launch_terminal_obs(const StepArgs &a, cudaStream_t st) {
    if (a.N <= 0 || !a.term_frames || !a.term_order) return cudaSuccess;
    terminal_obs_kernel<<<a.N, 128, 0, st>>>(a);
    return cudaGetLastError();
}

cudaError_t launch_rewarder(const RewarderArgs &a, cudaStream_t st) {
    if (a.n <= 0) return cudaSuccess;
    rewarder_kernel<<<(a.n + 3) / 4, 128, 0, st>>>(a);
    return cudaGetLastError();
}

}  // namespace ffmp
