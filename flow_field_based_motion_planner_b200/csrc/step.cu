// step.cu — SPEC.md §7/§8: the fused per-step kernel of the batched FFMP environment.
//
// One CTA (4 warps) per environment.  Warp 0 runs the scalar part — fp32 unicycle integration
// (SPEC K), relative goal / velocity (train.py:174-188), the 21-cell footprint collision test on the
// global flow image (ffmp.py:85-105), goal test, reward, done, truncation (ffmp.py:120-164,
// train.py:607-608), auto-reset onto the next pre-generated scenario slot — and broadcasts the crop
// origin through shared memory; all four warps then copy the ego-centred W x W window of the flow
// image into the observation frame ring with coalesced 32-bit stores (byte-realigned with a funnel
// shift).  HBM-bound: W^2 read + W^2 written per env-step (+ one more frame on ring wrap / reset).
#include "ffmp_kernels.cuh"

namespace ffmp {

namespace {

__constant__ int8_t FOOT_DI[21] = {-2, -2, -2, -1, -1, -1, -1, -1, 0, 0, 0, 0, 0, 1, 1, 1, 1, 1, 2, 2, 2};
__constant__ int8_t FOOT_DJ[21] = {-1, 0, 1, -2, -1, 0, 1, 2, -2, -1, 0, 1, 2, -2, -1, 0, 1, 2, -1, 0, 1};

struct Bcast {
    int active;          // this env writes frames in this call
    int two;             // also write the older frame
    int ci, cj;          // crop centre of the newest frame
    int pi, pj;          // crop centre of the older frame
    unsigned long long plane;  // scenario plane index of the flow image
};

__device__ __forceinline__ uint32_t crop_word(const uint8_t *__restrict__ img, int G, int i, int j0) {
    if (static_cast<unsigned>(i) >= static_cast<unsigned>(G)) return 0xFFFFFFFFu;
    const uint8_t *row = img + static_cast<size_t>(i) * G;
    if (j0 >= 0 && j0 + 3 < G) {
        const int m = j0 & 3;
        const uint32_t *wp = reinterpret_cast<const uint32_t *>(row + (j0 - m));
        const uint32_t lo = __ldg(wp);
        const uint32_t hi = m ? __ldg(wp + 1) : 0u;
        return __funnelshift_r(lo, hi, 8 * m);
    }
    uint32_t v = 0;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        const int j = j0 + c;
        const uint32_t b = static_cast<unsigned>(j) < static_cast<unsigned>(G) ? __ldg(row + j) : 255u;
        v |= b << (8 * c);
    }
    return v;
}

__global__ void __launch_bounds__(128) step_kernel(StepArgs a) {
    __shared__ Bcast bc;
    const int e = blockIdx.x;
    const int tid = threadIdx.x;
    const int G = a.G, W = a.W;
    const size_t cells = static_cast<size_t>(G) * G;

    if (tid < 32) {
        const int lane = tid;
        uint32_t *st = a.state + static_cast<size_t>(e) * ST_WORDS;
        float x = __uint_as_float(st[ST_X]), y = __uint_as_float(st[ST_Y]), yaw = __uint_as_float(st[ST_YAW]);
        const float gx = __uint_as_float(st[ST_GX]), gy = __uint_as_float(st[ST_GY]);
        const float d_first = __uint_as_float(st[ST_DFIRST]);
        float ep_return = __uint_as_float(st[ST_RETURN]);
        int steps = static_cast<int>(st[ST_STEPS]);
        uint32_t episode = st[ST_EPISODE];
        bool begin = false, active = true;
        int ci = 0, cj = 0, pi = 0, pj = 0;

        if (a.mode == 0) {
            long long act = a.actions[e];
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
            const float dx = fsub(gx, nx), dy = fsub(gy, ny);
            const float d = dist_spec(dx, dy);
            const float bearing = pi_to_pi(fsub(atan2_spec(dy, dx), nyaw));
            const float vl = dist_spec(fsub(nx, x), fsub(ny, y));
            const float va = pi_to_pi(fsub(nyaw, yaw));
            ci = robot_cell(nx);
            cj = robot_cell(ny);
            const uint8_t *img = a.flow + (static_cast<size_t>(episode % a.S) * a.N + e) * cells;
            bool hit = false;
            if (lane < 21) {
                const int i = ci + FOOT_DI[lane], j = cj + FOOT_DJ[lane];
                hit = static_cast<unsigned>(i) >= static_cast<unsigned>(G) || static_cast<unsigned>(j) >= static_cast<unsigned>(G);
                if (!hit) hit = __ldg(img + static_cast<size_t>(i) * G + j) == 255;
            }
            const bool col = __ballot_sync(FULL, hit) != 0;
            const bool goal = d < 0.5f;
            const float r = fadd(fadd(goal ? 1.0f : fmul(0.05f, fsub(d_first, d)), col ? -1.0f : 0.0f), -0.05f);
            steps += 1;
            const bool trunc = steps == a.max_steps;
            const bool done = col || goal || trunc;
            ep_return = fadd(ep_return, r);
            if (lane == 0) {
                a.reward[e] = r;
                a.done[e] = done ? 1 : 0;
                a.flags[e] = static_cast<uint8_t>((col ? 1 : 0) | (goal ? 2 : 0) | (trunc ? 4 : 0));
                a.term_rel_goal[2 * e] = d; a.term_rel_goal[2 * e + 1] = bearing;
                a.term_velocity[2 * e] = vl; a.term_velocity[2 * e + 1] = va;
            }
            if (done) {
                if (lane == 0) { a.fin_return[e] = ep_return; a.fin_length[e] = steps; }
                begin = true;
            } else {
                pi = robot_cell(x); pj = robot_cell(y);
                if (lane == 0) {
                    st[ST_X] = __float_as_uint(nx); st[ST_Y] = __float_as_uint(ny); st[ST_YAW] = __float_as_uint(nyaw);
                    st[ST_RETURN] = __float_as_uint(ep_return);
                    st[ST_STEPS] = static_cast<uint32_t>(steps);
                    a.rel_goal[2 * e] = d; a.rel_goal[2 * e + 1] = bearing;
                    a.velocity[2 * e] = vl; a.velocity[2 * e + 1] = va;
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
                    // the slot of the finished episode is refilled with episode + S - 1
                    const uint32_t idx = atomicAdd(a.regen_count, 1u);
                    a.regen_env[idx] = static_cast<uint32_t>(e);
                    a.regen_episode[idx] = episode + static_cast<uint32_t>(a.S) - 1u;
                }
            }
            const uint32_t *rec = a.scen + (static_cast<size_t>(episode % a.S) * a.N + e) * SC_WORDS;
            x = __uint_as_float(rec[SC_X0]); y = __uint_as_float(rec[SC_Y0]); yaw = __uint_as_float(rec[SC_YAW0]);
            const float ngx = __uint_as_float(rec[SC_GX]), ngy = __uint_as_float(rec[SC_GY]);
            const float dx = fsub(ngx, x), dy = fsub(ngy, y);
            const float d = dist_spec(dx, dy);
            const float bearing = pi_to_pi(fsub(atan2_spec(dy, dx), yaw));
            ci = pi = robot_cell(x);
            cj = pj = robot_cell(y);
            if (lane == 0) {
                st[ST_X] = __float_as_uint(x); st[ST_Y] = __float_as_uint(y); st[ST_YAW] = __float_as_uint(yaw);
                st[ST_GX] = __float_as_uint(ngx); st[ST_GY] = __float_as_uint(ngy);
                st[ST_DFIRST] = __float_as_uint(d);
                st[ST_RETURN] = __float_as_uint(0.0f);
                st[ST_STEPS] = 0u;
                st[ST_EPISODE] = episode;
                a.rel_goal[2 * e] = d; a.rel_goal[2 * e + 1] = bearing;
                a.velocity[2 * e] = 0.0f; a.velocity[2 * e + 1] = 0.0f;
                if (a.mode == 2) { a.reward[e] = 0.0f; a.done[e] = 0; a.flags[e] = 0; }
            }
        }
        if (lane == 0) {
            bc.active = active ? 1 : 0;
            bc.two = (begin || a.write_older) ? 1 : 0;
            bc.ci = ci; bc.cj = cj; bc.pi = pi; bc.pj = pj;
            bc.plane = static_cast<unsigned long long>(episode % a.S) * a.N + e;
        }
    }
    __syncthreads();
    if (!bc.active) return;

    // ---- observation: crop the flow image into the frame ring ------------------------------------
    const uint8_t *img = a.flow + static_cast<size_t>(bc.plane) * cells;
    const int wpr = W >> 2;
    const int nwords = W * wpr;
    uint32_t *f_new = reinterpret_cast<uint32_t *>(a.frames + (static_cast<size_t>(e) * a.K + a.slot_new) * W * W);
    uint32_t *f_old = reinterpret_cast<uint32_t *>(a.frames + (static_cast<size_t>(e) * a.K + a.slot_new - 1) * W * W);
    const int i0 = bc.ci - (W >> 1), j00 = bc.cj - (W >> 1);
    const bool two = bc.two != 0;
    const bool same = bc.pi == bc.ci && bc.pj == bc.cj;
    const int da = 128 / wpr, db = 128 - da * wpr;
    int row = tid / wpr, bw = tid - row * wpr;
    if (two && !same) {
        const int p0 = bc.pi - (W >> 1), q0 = bc.pj - (W >> 1);
        for (int q = tid; q < nwords; q += 128) {
            f_new[q] = crop_word(img, G, i0 + row, j00 + 4 * bw);
            f_old[q] = crop_word(img, G, p0 + row, q0 + 4 * bw);
            row += da; bw += db;
            if (bw >= wpr) { bw -= wpr; row += 1; }
        }
    } else {
        for (int q = tid; q < nwords; q += 128) {
            const uint32_t v = crop_word(img, G, i0 + row, j00 + 4 * bw);
            f_new[q] = v;
            if (two) f_old[q] = v;
            row += da; bw += db;
            if (bw >= wpr) { bw -= wpr; row += 1; }
        }
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
            const float r = a.scan[static_cast<size_t>(item) * a.scan_len + k];
            // `if scan[i]: if scan[i] < 0.13` with None encoded as NaN; the threshold is the fp64 0.13
            if (r == r && r != 0.0f && static_cast<double>(r) < 0.13) hit = true;
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

cudaError_t launch_step(const StepArgs &a, cudaStream_t st) {
    if (a.N <= 0) return cudaSuccess;
    step_kernel<<<a.N, 128, 0, st>>>(a);
    return cudaGetLastError();
}

cudaError_t launch_rewarder(const RewarderArgs &a, cudaStream_t st) {
    if (a.n <= 0) return cudaSuccess;
    rewarder_kernel<<<(a.n + 3) / 4, 128, 0, st>>>(a);
    return cudaGetLastError();
}

}  // namespace ffmp
