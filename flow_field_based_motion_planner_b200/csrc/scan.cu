// scan.cu — SPEC.md §9: LiDAR scan synthesis, the input of FFMP.rewarder2 / is_collision2
// (/root/reference/src/gym_ffmp/envs/ffmp.py:108-117,179-188; call site src/train.py:577; the reference took the scan from
// Gazebo's /scan topic, src/train.py:87,144-150).  SURVEY.md §8(f) row 3.
//
// One CTA per pose, one thread per beam: an exact grid traversal (Amanatides-Woo) with one fp32 rounding per written
// operation, so ranges are bit-identical to the oracle.  The cells a scan touches lie within RMAX of the robot
// (<= 141 x 141 bytes at 3.5 m): after the first beams they are L1 hits, so the kernel is bound by the dependent
// compare / select / load chain of the traversal (~45 cell steps per beam on the bench maps), not by HBM.
#include <cmath>
#include <cstdlib>

#include "ffmp_kernels.cuh"

namespace ffmp {

namespace {

template <bool FLOW>
__device__ __forceinline__ bool scan_blocked(const uint8_t *__restrict__ map, int G, int i, int j) {
    if (static_cast<unsigned>(i) >= static_cast<unsigned>(G) || static_cast<unsigned>(j) >= static_cast<unsigned>(G)) return true;
    const uint8_t v = __ldg(map + i * G + j);
    return FLOW ? v == 255 : v != 0;
}

// Pose and plane of item e: the operator reads pose[e] and plane e, the env reads the state record and the current slot.
__device__ __forceinline__ void scan_item(const ScanArgs &a, int e, float &x, float &y, float &yaw, size_t &plane) {
    plane = static_cast<size_t>(e);
    if (a.pose) {
        x = a.pose[3 * e]; y = a.pose[3 * e + 1]; yaw = a.pose[3 * e + 2];
    } else {
        const uint32_t *st = a.state + static_cast<size_t>(e) * ST_WORDS;
        x = __uint_as_float(st[ST_X]); y = __uint_as_float(st[ST_Y]); yaw = __uint_as_float(st[ST_YAW]);
        plane = static_cast<size_t>(st[ST_EPISODE] % static_cast<uint32_t>(a.S)) * a.N + e;
    }
}

struct BeamSetup {
    float tdx, tmx, tdy, tmy;
    int si, sj;
};

// SPEC.md §9 per-beam constants; fu / fv = fractional position of the robot inside its cell
__device__ __forceinline__ BeamSetup beam_setup(float yaw, int k, float inc, float fu, float fv) {
    const float inf = __int_as_float(0x7F800000);
    float s, c;
    sincos_spec(pi_to_pi(fadd(yaw, fmul(static_cast<float>(k), inc))), s, c);
    BeamSetup b;
    if (c > 0.0f) { b.si = 1; b.tdx = __fdiv_rn(1.0f, c); b.tmx = fmul(fsub(1.0f, fu), b.tdx); }
    else if (c < 0.0f) { b.si = -1; b.tdx = __fdiv_rn(1.0f, -c); b.tmx = fmul(fu, b.tdx); }
    else { b.si = 0; b.tdx = b.tmx = inf; }
    if (s > 0.0f) { b.sj = 1; b.tdy = __fdiv_rn(1.0f, s); b.tmy = fmul(fsub(1.0f, fv), b.tdy); }
    else if (s < 0.0f) { b.sj = -1; b.tdy = __fdiv_rn(1.0f, -s); b.tmy = fmul(fv, b.tdy); }
    else { b.sj = 0; b.tdy = b.tmy = inf; }
    return b;
}

// Generic variant (any range_max): every cell test is a global byte load.
template <bool FLOW>
__global__ void __launch_bounds__(128) scan_kernel(ScanArgs a) {
    const int e = blockIdx.x;
    const int G = a.G;
    float x, y, yaw;
    size_t plane;
    scan_item(a, e, x, y, yaw, plane);
    const uint8_t *map = a.map + plane * (static_cast<size_t>(G) * G);
    float *out = a.scan + static_cast<size_t>(e) * a.beams;
    const float inc = __fdiv_rn(TWO_PI_F, static_cast<float>(a.beams));
    const float u0 = fadd(fmul(x, INV_RES), 0.5f), v0 = fadd(fmul(y, INV_RES), 0.5f);
    const float fl_u = floorf(u0), fl_v = floorf(v0);
    const int i0 = static_cast<int>(fl_u), j0 = static_cast<int>(fl_v);
    bool hit = false;
    if (scan_blocked<FLOW>(map, G, i0, j0)) {
        for (int k = threadIdx.x; k < a.beams; k += blockDim.x) out[k] = 0.0f;
    } else {
        const float fu = fsub(u0, fl_u), fv = fsub(v0, fl_v);
        const float max_t = fmul(a.range_max, INV_RES);
        const float inf = __int_as_float(0x7F800000);
        for (int k = threadIdx.x; k < a.beams; k += blockDim.x) {
            BeamSetup b = beam_setup(yaw, k, inc, fu, fv);
            int i = i0, j = j0;
            float r;
            for (;;) {
                float t;
                if (b.tmx < b.tmy) { t = b.tmx; i += b.si; b.tmx = fadd(b.tmx, b.tdx); }
                else { t = b.tmy; j += b.sj; b.tmy = fadd(b.tmy, b.tdy); }
                if (t > max_t) { r = inf; break; }
                if (scan_blocked<FLOW>(map, G, i, j)) { r = fmul(t, RES); break; }
            }
            out[k] = r;
            hit |= r != 0.0f && static_cast<double>(r) < 0.13;     // ffmp.py:108-117 compares in fp64
        }
    }
    if (a.hit) {
        const int any = __syncthreads_or(hit ? 1 : 0);
        if (threadIdx.x == 0) a.hit[e] = any ? 1 : 0;
    }
}

// Windowed variant: the cells within reach of the sensor (|di|, |dj| <= R = ceil(RMAX / RES) + 2 around the robot cell)
// are packed once per pose into a "blocked" bit window in shared memory (256 bits per row, so a cell is ONE integer
// pos = row * 256 + col and a traversal step is one select + add on it); out-of-grid cells are blocked.  The global byte
// loads of the generic variant touch up to 32 different lines per warp instruction (the beams of a warp fan out over
// different rows): the LSU then serialises ~20 wavefronts per cell test, which bounded the generic kernel.
constexpr int SCAN_WIN_MAX = 224;      // window side limit of this variant: 7 words of 32 columns per row
constexpr int SCAN_ROW_BITS = 256;

template <bool FLOW>
__global__ void __launch_bounds__(128) scan_window_kernel(ScanArgs a, int R) {
    extern __shared__ uint32_t win[];              // [2R+1][8] words
    const int e = blockIdx.x;
    const int G = a.G;
    float x, y, yaw;
    size_t plane;
    scan_item(a, e, x, y, yaw, plane);
    const uint8_t *map = a.map + plane * (static_cast<size_t>(G) * G);
    float *out = a.scan + static_cast<size_t>(e) * a.beams;
    const float inc = __fdiv_rn(TWO_PI_F, static_cast<float>(a.beams));
    const float u0 = fadd(fmul(x, INV_RES), 0.5f), v0 = fadd(fmul(y, INV_RES), 0.5f);
    const float fl_u = floorf(u0), fl_v = floorf(v0);
    // poses far outside the grid behave like "robot cell blocked" (clamped so that the integer conversion is defined)
    const int i0 = static_cast<int>(fminf(fmaxf(fl_u, -4.0f), static_cast<float>(G) + 4.0f));
    const int j0 = static_cast<int>(fminf(fmaxf(fl_v, -4.0f), static_cast<float>(G) + 4.0f));
    const int rows = 2 * R + 1;
    const int wi0 = i0 - R;
    const int wj0 = (j0 - R) & ~3;                 // 4-byte aligned window origin (G % 4 == 0): every 32-bit load is in or out
    const int wpr = (j0 + R - wj0) / 32 + 1;       // words per row that hold window columns (<= 7)

    // ---- pack: one thread per (row, word); 8 aligned 32-bit loads -> 32 blocked bits ------------------------------
    for (int idx = threadIdx.x; idx < rows * wpr; idx += blockDim.x) {
        const int li = idx / wpr, w = idx - li * wpr;
        const int gi = wi0 + li, gj = wj0 + 32 * w;
        uint32_t bits = 0xFFFFFFFFu;
        if (static_cast<unsigned>(gi) < static_cast<unsigned>(G)) {
            const uint32_t *src = reinterpret_cast<const uint32_t *>(map + static_cast<size_t>(gi) * G) + (gj >> 2);   // gj % 4 == 0, may be negative
            uint32_t v[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const int c = gj + 4 * q;
                v[q] = static_cast<unsigned>(c) < static_cast<unsigned>(G) ? __ldg(src + q) : 0xFFFFFFFFu;
            }
            bits = 0;
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                uint32_t m;
                if (FLOW) m = v[q] & 0x01010101u;           // 255 is the only odd value of a flow image (direction codes are 28 k)
                else m = ((((v[q] & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | v[q]) >> 7) & 0x01010101u;   // byte != 0
                bits |= ((m * 0x10204080u) >> 28) << (4 * q);   // bytes 0..3 -> bits 0..3
            }
        }
        win[li * (SCAN_ROW_BITS / 32) + w] = bits;
    }
    __syncthreads();

    auto blocked = [&](int pos) -> bool {
        return (__funnelshift_r(win[pos >> 5], 0u, pos) & 1u) != 0;     // shift count wraps at 32
    };
    const int pos0 = R * SCAN_ROW_BITS + (j0 - wj0);
    bool hit = false;
    if (blocked(pos0)) {
        for (int k = threadIdx.x; k < a.beams; k += blockDim.x) out[k] = 0.0f;
    } else {
        const float fu = fsub(u0, fl_u), fv = fsub(v0, fl_v);
        const float max_t = fmul(a.range_max, INV_RES);
        const float inf = __int_as_float(0x7F800000);
        for (int k = threadIdx.x; k < a.beams; k += blockDim.x) {
            const BeamSetup b = beam_setup(yaw, k, inc, fu, fv);
            float tmx = b.tmx, tmy = b.tmy;
            const int dpx = b.si * SCAN_ROW_BITS, dpy = b.sj;
            int pos = pos0;
            float r;
            for (;;) {
                // one step of SPEC §9: the smaller crossing parameter wins, ties step in j; x + 0.0f == x exactly
                const bool px = tmx < tmy;
                const float t = px ? tmx : tmy;
                tmx = fadd(tmx, px ? b.tdx : 0.0f);
                tmy = fadd(tmy, px ? 0.0f : b.tdy);
                pos += px ? dpx : dpy;
                if (t > max_t) { r = inf; break; }
                if (blocked(pos)) { r = fmul(t, RES); break; }
            }
            out[k] = r;
            hit |= r != 0.0f && static_cast<double>(r) < 0.13;     // ffmp.py:108-117 compares in fp64
        }
    }
    if (a.hit) {
        const int any = __syncthreads_or(hit ? 1 : 0);
        if (threadIdx.x == 0) a.hit[e] = any ? 1 : 0;
    }
}

}  // namespace

cudaError_t launch_scan(const ScanArgs &a, cudaStream_t st) {
    if (a.n <= 0 || a.beams <= 0) return cudaSuccess;
    // window radius: a tested cell has t <= RMAX / RES, so it lies within ceil(RMAX / RES) + 1 cells of the robot cell
    const double cells = std::ceil(static_cast<double>(a.range_max) * 20.0) + 2.0;
    const bool windowed = a.G % 4 == 0 && reinterpret_cast<uintptr_t>(a.map) % 4 == 0 && cells * 2 + 1 + 3 <= SCAN_WIN_MAX && std::getenv("FFMP_SCAN_GENERIC") == nullptr;
    if (windowed) {
        const int R = static_cast<int>(cells);
        const size_t smem = static_cast<size_t>(2 * R + 1) * (SCAN_ROW_BITS / 32) * sizeof(uint32_t);
        if (a.flow_mode) scan_window_kernel<true><<<a.n, 128, smem, st>>>(a, R);
        else scan_window_kernel<false><<<a.n, 128, smem, st>>>(a, R);
    } else if (a.flow_mode) {
        scan_kernel<true><<<a.n, 128, 0, st>>>(a);
    } else {
        scan_kernel<false><<<a.n, 128, 0, st>>>(a);
    }
    return cudaGetLastError();
}

}  // namespace ffmp
