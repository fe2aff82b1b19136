// dmf_kernels.cuh — sm_100a kernels of the dense monocular depth filter.
//
// Replaces, for every interior pixel of the reference frame and each new frame, the CPU
// loop of luigifreda/slamplay dense_mapping/test_monocular_mapping.cpp ("ref:LINE"):
//   update ref:355-393, epipolarSearch ref:397-447, NCC ref:449-480,
//   getBilinearInterpolatedValue ref:165-174, updateDepthFilter ref:482-567.
//
// Design (DESIGN.md §3 has the full derivation):
//  * One CTA owns a TILE_W x TILE_H tile of reference pixels and runs three phases in ONE
//    kernel, so depth / depth_cov2 are read once and written once per frame in HBM:
//      P1 (thread = pixel, FP64): gate ref:366, projections of mu and mu±3σ ref:402-422,
//         sample count of the l-loop ref:432; packs the pixel's 7x7 reference patch.
//      P2 (thread = (pixel, sample), block-local flattened work list built by a prefix sum
//         over the sample counts): one NCC per item.  Lanes are always full, whatever the
//         per-pixel search length (0..286 samples).
//      P3 (thread = pixel, FP64): triangulation + uncertainty + Gaussian fusion ref:482-567.
//  * NCC arithmetic.  All 49 taps of one NCC share the same four bilinear weights (the tap
//    offsets are integers, ref:461), so every sum the ZNCC needs is a linear / quadratic form
//    in (w00,w10,w01,w11) over INTEGER moments of the 8x8 u8 block under the sample:
//    4 window sums S, 4 cross sums R with the reference patch, 10 Gram sums G.  They are
//    accumulated exactly with IDP.4A (4 u8 MACs per instruction, 192 per NCC), centred exactly
//    in int32 (49*R - Sr*S, 49*G - S*S'), and only the final 30-flop combination is FP32.
//    That is the reference's two-pass (centred) ZNCC up to ~1e-7, with no u8->f32 conversion
//    in the loop and none of the cancellation of a one-pass FP32 variance.
//  * Current-image taps come from global memory through aligned 32-bit __ldg gathers plus
//    funnel shifts (texture units filter with 8-bit weights and would break parity).
//  * arg-max keeps the reference's "first strict maximum" (ref:438) through a 64-bit key
//    (ordered NCC bits : 0xFFFFFFFE - sample index; 0xFFFFFFFF is the "no winner yet" sentinel that
//    goes with best_ncc = -1.0) and shared-memory atomicMax.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace dmf {

constexpr int TILE_W = 32;
constexpr int TILE_H = 8;
constexpr int TILE_PIX = TILE_W * TILE_H;  // == threads per CTA
constexpr int NCC_AREA = 49;
// 1e-10 * (49*255^2)^2 : the reference's epsilon (ref:479) in centred-integer units
constexpr float NCC_EPS_INT = 1015.2029750625f;

struct KParams {
    int width, height, border;
    int row_begin, row_end;  // interior rows owned by this context
    int inverse_depth;
    int write_flags;
    float ncc_thresh;
    double fx, fy, cx, cy;
    double step, max_half_len, min_depth, n_sigma, min_cov, max_cov;
    double q[4], t[3];    // T_C_R (unit quaternion x,y,z,w + translation)
    double qi[4], ti[3];  // T_R_C = T_C_R^-1 (ref:491), computed on the host
    const uint8_t *curr;  // pitched, 4-byte aligned rows
    const uint8_t *ref;
    const int2 *refstat;  // per pixel: (sum r, 49*sum r^2 - (sum r)^2)
    double *depth;
    double *cov2;
    uint8_t *flags;
    float *dbg_ncc;  // with write_flags: best NCC per active pixel
    int *dbg_n;      // with write_flags: (samples << 16) | winning sample index (0xFFFF: none)
    unsigned long long *counters;  // [0]=active [1]=ncc_evals [2]=accepted
    int curr_pitch, ref_pitch, stat_pitch, state_pitch, flags_pitch;  // in elements
};

// ----------------------------------------------------------------------------------------
// small FP64 helpers
struct D3 { double x, y, z; };
__device__ __forceinline__ double dot3(const D3 &a, const D3 &b) { return a.x * b.x + (a.y * b.y + a.z * b.z); }
__device__ __forceinline__ D3 cross3(const D3 &a, const D3 &b) {
    return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
// Eigen Quaternion::_transformVector as used by Sophus SE3 * point
__device__ __forceinline__ D3 qrot(const double q[4], const D3 &v) {
    D3 qv{q[0], q[1], q[2]};
    D3 uv = cross3(qv, v);
    uv.x += uv.x; uv.y += uv.y; uv.z += uv.z;
    D3 c = cross3(qv, uv);
    return {v.x + q[3] * uv.x + c.x, v.y + q[3] * uv.y + c.y, v.z + q[3] * uv.z + c.z};
}
__device__ __forceinline__ void normalize3(D3 &a) {  // Eigen normalize(): only when squaredNorm > 0
    double z = dot3(a, a);
    if (z > 0) { double n = sqrt(z); a.x /= n; a.y /= n; a.z /= n; }
}
// exact int -> double for |k| < 2^31 without the slow I2F.F64 path
__device__ __forceinline__ double int2double_fast(int k) {
    return __hiloint2double(0x43300000, (int)((unsigned)k ^ 0x80000000u)) - 4503601774854144.0;  // 2^52 + 2^31
}

// sample position parameter of iteration k of the loop ref:432, l_k = -half + step*k
__device__ __forceinline__ double sample_l(double half, double step, int k) {
    return fma(step, int2double_fast(k), -half);
}

// ----------------------------------------------------------------------------------------
// K1: once per reference frame — reference-patch statistics (ref half of NCC, ref:458-459,468,476)
// stat.x = sum of the 49 bytes, stat.y = 49*sum(b^2) - (sum b)^2   (both exact in int32)
__global__ void __launch_bounds__(256) ref_stats_kernel(const uint8_t *__restrict__ ref, int ref_pitch, int width,
                                                        int height, int border, int2 *__restrict__ stat,
                                                        int stat_pitch) {
    int x = blockIdx.x * blockDim.x + threadIdx.x + border;
    int y = blockIdx.y * blockDim.y + threadIdx.y + border;
    if (x >= width - border || y >= height - border) return;
    int s = 0, s2 = 0;
#pragma unroll
    for (int dy = -3; dy <= 3; ++dy) {
        const uint8_t *row = ref + (size_t)(y + dy) * ref_pitch + (x - 3);
#pragma unroll
        for (int dx = 0; dx < 7; ++dx) {
            int b = row[dx];
            s += b;
            s2 += b * b;
        }
    }
    stat[(size_t)y * stat_pitch + x] = make_int2(s, NCC_AREA * s2 - s * s);
}

// ----------------------------------------------------------------------------------------
// Loads 8 bytes starting at an arbitrary byte address from 4-byte aligned words.
__device__ __forceinline__ void load_row8(const uint32_t *wp, unsigned sh, uint32_t &lo, uint32_t &hi) {
    uint32_t w0 = __ldg(wp), w1 = __ldg(wp + 1), w2 = __ldg(wp + 2);
    lo = __funnelshift_r(w0, w1, sh);
    hi = __funnelshift_r(w1, w2, sh);
}

__device__ __forceinline__ int dp4(uint32_t a, uint32_t b, int c) { return (int)__dp4a(a, b, (unsigned)c); }

// Per-pixel shared-memory record written by P1 and read by P2/P3.
struct __align__(16) RefPatch {
    uint32_t row[14];  // rows dy=-3..3: (lo = bytes dx -3..0, hi = bytes dx 1..3 and a zero byte)
    int sum;           // Sr
    int den1;          // 49*sum r^2 - Sr^2
};

struct Shared {
    RefPatch patch[TILE_PIX];
    double pmx[TILE_PIX], pmy[TILE_PIX];  // px_mean_curr ref:406
    double dx[TILE_PIX], dy[TILE_PIX];    // epipolar_direction ref:419-420
    double half[TILE_PIX];                // half_length ref:421-422
    unsigned long long best[TILE_PIX];    // arg-max key
    int offs[TILE_PIX + 1];               // exclusive prefix sum of the sample counts
    int warp_sum[TILE_PIX / 32];
    unsigned int cnt_active, cnt_eval, cnt_accept;
};

__device__ __forceinline__ unsigned int ordered_bits(float f) {
    unsigned int u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float from_ordered_bits(unsigned int u) {
    return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}

// One NCC (ref:449-480) of reference pixel record `rp` against the current image at the
// sub-pixel position whose integer part is (ix,iy) and bilinear fractions (fx,fy).
__device__ __forceinline__ float ncc_int_moments(const KParams &P, const RefPatch &rp, int ix, int iy, float fx,
                                                 float fy) {
    const uint8_t *base = P.curr + (size_t)(iy - 3) * P.curr_pitch + (ix - 3);
    unsigned mis = (unsigned)(reinterpret_cast<uintptr_t>(base) & 3u);
    const uint32_t *wp = reinterpret_cast<const uint32_t *>(base - mis);
    const unsigned sh = mis * 8u;
    const int pitch_w = P.curr_pitch >> 2;
    const uint32_t ONES = 0x01010101u;

    // reference rows (16-byte shared loads)
    const uint4 *rq = reinterpret_cast<const uint4 *>(rp.row);
    uint4 r0 = rq[0], r1 = rq[1], r2 = rq[2], r3 = rq[3];
    const uint32_t Rlo[7] = {r0.x, r0.z, r1.x, r1.z, r2.x, r2.z, r3.x};
    const uint32_t Rhi[7] = {r0.y, r0.w, r1.y, r1.w, r2.y, r2.w, r3.y};
    const int Sr = (int)r3.z, den1 = (int)r3.w;

    // all 24 gathers first (memory-level parallelism), then the integer moments
    uint32_t lo[8], hi[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) load_row8(wp + j * pitch_w, sh, lo[j], hi[j]);

    // accumulators: top = block row 0, mid = rows 1..6, bot = row 7
    int s0t = 0, s0m = 0, s0b = 0, s1t = 0, s1m = 0, s1b = 0;  // row sums, column window a=0 / a=1
    int q0t = 0, q0m = 0, q0b = 0, q1t = 0, q1m = 0, q1b = 0;  // sum of squares
    int ht = 0, hm = 0, hb = 0;                                // horizontal neighbour products
    int v0 = 0, v1 = 0, d01 = 0, d10 = 0;                      // vertical / diagonal products (rows j, j+1)
    int R00 = 0, R10 = 0, R01 = 0, R11 = 0;                    // cross sums with the reference patch
    uint32_t p0l = 0, p0h = 0, p1l = 0, p1h = 0;               // previous row
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const uint32_t x0l = lo[j], x0h = hi[j] & 0x00FFFFFFu;                       // columns 0..6
        const uint32_t x1l = __funnelshift_r(lo[j], hi[j], 8), x1h = hi[j] >> 8;     // columns 1..7
        int s0 = dp4(x0l, ONES, dp4(x0h, ONES, 0));
        int s1 = dp4(x1l, ONES, dp4(x1h, ONES, 0));
        int q0 = dp4(x0l, x0l, dp4(x0h, x0h, 0));
        int q1 = dp4(x1l, x1l, dp4(x1h, x1h, 0));
        int h = dp4(x0l, x1l, dp4(x0h, x1h, 0));
        if (j == 0) { s0t = s0; s1t = s1; q0t = q0; q1t = q1; ht = h; }
        else if (j == 7) { s0b = s0; s1b = s1; q0b = q0; q1b = q1; hb = h; }
        else { s0m += s0; s1m += s1; q0m += q0; q1m += q1; hm += h; }
        if (j > 0) {
            v0 = dp4(p0l, x0l, dp4(p0h, x0h, v0));
            v1 = dp4(p1l, x1l, dp4(p1h, x1h, v1));
            d01 = dp4(p0l, x1l, dp4(p0h, x1h, d01));
            d10 = dp4(p1l, x0l, dp4(p1h, x0h, d10));
            R01 = dp4(Rlo[j - 1], x0l, dp4(Rhi[j - 1], x0h, R01));
            R11 = dp4(Rlo[j - 1], x1l, dp4(Rhi[j - 1], x1h, R11));
        }
        if (j < 7) {
            R00 = dp4(Rlo[j], x0l, dp4(Rhi[j], x0h, R00));
            R10 = dp4(Rlo[j], x1l, dp4(Rhi[j], x1h, R10));
        }
        p0l = x0l; p0h = x0h; p1l = x1l; p1h = x1h;
    }
    // window (a,b): columns a..a+6, rows b..b+6
    const int S00 = s0t + s0m, S01 = s0m + s0b, S10 = s1t + s1m, S11 = s1m + s1b;
    // exact centring in int32 (all terms < 2^31)
    const int cR00 = NCC_AREA * R00 - Sr * S00, cR10 = NCC_AREA * R10 - Sr * S10;
    const int cR01 = NCC_AREA * R01 - Sr * S01, cR11 = NCC_AREA * R11 - Sr * S11;
    const int G0000 = NCC_AREA * (q0t + q0m) - S00 * S00, G0101 = NCC_AREA * (q0m + q0b) - S01 * S01;
    const int G1010 = NCC_AREA * (q1t + q1m) - S10 * S10, G1111 = NCC_AREA * (q1m + q1b) - S11 * S11;
    const int G0010 = NCC_AREA * (ht + hm) - S00 * S10, G0111 = NCC_AREA * (hm + hb) - S01 * S11;
    const int G0001 = NCC_AREA * v0 - S00 * S01, G1011 = NCC_AREA * v1 - S10 * S11;
    const int G0011 = NCC_AREA * d01 - S00 * S11, G1001 = NCC_AREA * d10 - S10 * S01;

    // FP32 combination with the bilinear weights of ref:169-172
    const float gx = 1.0f - fx, gy = 1.0f - fy;
    const float w00 = gx * gy, w10 = fx * gy, w01 = gx * fy, w11 = fx * fy;
    float num = w00 * (float)cR00;
    num = fmaf(w10, (float)cR10, num);
    num = fmaf(w01, (float)cR01, num);
    num = fmaf(w11, (float)cR11, num);
    // den2 = w^T G w
    float a0 = w00 * (float)G0000;
    a0 = fmaf(w10, (float)G0010, a0); a0 = fmaf(w01, (float)G0001, a0); a0 = fmaf(w11, (float)G0011, a0);
    float a1 = w00 * (float)G0010;
    a1 = fmaf(w10, (float)G1010, a1); a1 = fmaf(w01, (float)G1001, a1); a1 = fmaf(w11, (float)G1011, a1);
    float a2 = w00 * (float)G0001;
    a2 = fmaf(w10, (float)G1001, a2); a2 = fmaf(w01, (float)G0101, a2); a2 = fmaf(w11, (float)G0111, a2);
    float a3 = w00 * (float)G0011;
    a3 = fmaf(w10, (float)G1011, a3); a3 = fmaf(w01, (float)G0111, a3); a3 = fmaf(w11, (float)G1111, a3);
    float den2 = w00 * a0;
    den2 = fmaf(w10, a1, den2); den2 = fmaf(w01, a2, den2); den2 = fmaf(w11, a3, den2);
    den2 = fmaxf(den2, 0.0f);
    const float dd = fmaf((float)den1, den2, NCC_EPS_INT);
    float r = rsqrtf(dd);
    r = r * fmaf(-0.5f * dd * r, r, 1.5f);  // one Newton step: MUFU.RSQ is only ~2 ulp
    return num * r;
}

// ----------------------------------------------------------------------------------------
// K2: the fused per-frame update.
__global__ void __launch_bounds__(TILE_PIX, 3) update_fused_kernel(const __grid_constant__ KParams P) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Shared &S = *reinterpret_cast<Shared *>(smem_raw);
    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    const int x = P.border + blockIdx.x * TILE_W + (tid % TILE_W);
    const int y = P.row_begin + blockIdx.y * TILE_H + (tid / TILE_W);
    const bool in_img = (x < P.width - P.border) && (y < P.row_end);

    if (tid == 0) { S.cnt_active = 0; S.cnt_eval = 0; S.cnt_accept = 0; }

    // ------------------------------------------------------------------ P1
    double mu = 0, c2 = 0;
    D3 f_ref{0, 0, 1};
    int n = 0;
    bool active = false;
    if (in_img) {
        c2 = P.cov2[(size_t)y * P.state_pitch + x];
        active = !(c2 < P.min_cov || c2 > P.max_cov);  // ref:366 — NaN passes the gate
    }
    if (active) {
        mu = P.depth[(size_t)y * P.state_pitch + x];
        const double sigma = sqrt(c2);  // ref:377
        f_ref = D3{((double)x - P.cx) / P.fx, ((double)y - P.cy) / P.fy, 1.0};  // ref:207-212
        normalize3(f_ref);
        const D3 Rf = qrot(P.q, f_ref);  // T*(f*d) = d*(R f) + t
        double d_min, d_max;
        if (P.inverse_depth) {  // ref:407-410
            const double inv_mu = 1.0 / mu;
            d_min = 1.0 / (inv_mu + P.n_sigma * sigma);
            d_max = 1.0 / (inv_mu - P.n_sigma * sigma);
        } else {  // ref:412
            d_min = mu - P.n_sigma * sigma;
            d_max = mu + P.n_sigma * sigma;
        }
        if (d_min < P.min_depth) d_min = P.min_depth;  // ref:414
        // cam2px ref:215-219 of the three points
        const double zm = fma(Rf.z, mu, P.t[2]), z0 = fma(Rf.z, d_min, P.t[2]), z1 = fma(Rf.z, d_max, P.t[2]);
        const double pmx = fma(Rf.x, mu, P.t[0]) * P.fx / zm + P.cx, pmy = fma(Rf.y, mu, P.t[1]) * P.fy / zm + P.cy;
        const double p0x = fma(Rf.x, d_min, P.t[0]) * P.fx / z0 + P.cx, p0y = fma(Rf.y, d_min, P.t[1]) * P.fy / z0 + P.cy;
        const double p1x = fma(Rf.x, d_max, P.t[0]) * P.fx / z1 + P.cx, p1y = fma(Rf.y, d_max, P.t[1]) * P.fy / z1 + P.cy;
        double lx = p1x - p0x, ly = p1y - p0y;  // ref:418
        const double len2 = lx * lx + ly * ly;
        const double len = sqrt(len2);
        double half = 0.5 * len;  // ref:421
        if (len2 > 0) { lx /= len; ly /= len; }  // ref:420 (guarded normalize)
        if (half > P.max_half_len) half = P.max_half_len;  // ref:422
        // trip count of `for (l = -half; l <= half; l += step)` ref:432 (NaN half -> 0)
        if (half >= 0) {
            n = (int)(2.0 * half / P.step) + 1;
            while (n > 0 && sample_l(half, P.step, n - 1) > half) --n;
            while (n < 100000 && sample_l(half, P.step, n) <= half) ++n;
        }
        S.pmx[tid] = pmx; S.pmy[tid] = pmy; S.dx[tid] = lx; S.dy[tid] = ly; S.half[tid] = half;
        // pack the 7x7 reference patch of (x,y)
        const uint8_t *rb = P.ref + (size_t)(y - 3) * P.ref_pitch + (x - 3);
        unsigned mis = (unsigned)(reinterpret_cast<uintptr_t>(rb) & 3u);
        const uint32_t *wp = reinterpret_cast<const uint32_t *>(rb - mis);
        const int pw = P.ref_pitch >> 2;
#pragma unroll
        for (int j = 0; j < 7; ++j) {
            uint32_t lo, hi;
            load_row8(wp + j * pw, mis * 8u, lo, hi);
            S.patch[tid].row[2 * j] = lo;
            S.patch[tid].row[2 * j + 1] = hi & 0x00FFFFFFu;
        }
        const int2 st = P.refstat[(size_t)y * P.stat_pitch + x];
        S.patch[tid].sum = st.x;
        S.patch[tid].den1 = st.y;
    }
    S.best[tid] = ((unsigned long long)ordered_bits(-1.0f) << 32) | 0xFFFFFFFFull;  // best_ncc = -1.0 ref:430

    // block-wide exclusive prefix sum of n
    int incl = n;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += v;
    }
    if (lane == 31) S.warp_sum[warp] = incl;
    __syncthreads();
    int wbase = 0;
#pragma unroll
    for (int w = 0; w < TILE_PIX / 32; ++w) wbase += (w < warp) ? S.warp_sum[w] : 0;
    S.offs[tid] = wbase + incl - n;
    if (tid == TILE_PIX - 1) S.offs[TILE_PIX] = wbase + incl;
    __syncthreads();
    const int total = S.offs[TILE_PIX];

    // ------------------------------------------------------------------ P2
    unsigned int my_evals = 0;
    for (int item = tid; item < total; item += TILE_PIX) {
        // pixel of this item: largest p with offs[p] <= item
        int lo = 0, hi = TILE_PIX - 1;
#pragma unroll
        for (int it = 0; it < 8; ++it) {
            int mid = (lo + hi + 1) >> 1;
            if (S.offs[mid] <= item) lo = mid; else hi = mid - 1;
        }
        const int p = lo;
        const int k = item - S.offs[p];
        const double l = sample_l(S.half[p], P.step, k);
        const double sx = fma(l, S.dx[p], S.pmx[p]);  // ref:433
        const double sy = fma(l, S.dy[p], S.pmy[p]);
        // inside() ref:222-224
        const bool ok = sx >= P.border && sy >= P.border && sx + P.border < P.width && sy + P.border <= P.height;
        if (!ok) continue;
        const int ix = (int)sx, iy = (int)sy;  // positive: trunc == floor
        const float fx = (float)(sx - (double)ix), fy = (float)(sy - (double)iy);
        const float v = ncc_int_moments(P, S.patch[p], ix, iy, fx, fy);
        ++my_evals;
        const unsigned long long key = ((unsigned long long)ordered_bits(v) << 32) | (unsigned long long)(0xFFFFFFFEu - (unsigned)k);
        if (v == v && key > S.best[p]) atomicMax(&S.best[p], key);  // first strict maximum ref:438-441
    }
    __syncthreads();

    // ------------------------------------------------------------------ P3
    bool accepted = false;
    if (active) {
        const unsigned long long key = S.best[tid];
        const float best = from_ordered_bits((unsigned)(key >> 32));
        accepted = !(best < P.ncc_thresh);  // ref:443
        if ((unsigned)key == 0xFFFFFFFFu) accepted = false;  // no sample beat -1.0
    }
    if (accepted) {
        const int k = (int)(0xFFFFFFFEu - (unsigned)S.best[tid]);
        const double l = sample_l(S.half[tid], P.step, k);
        const double ex = S.dx[tid], ey = S.dy[tid];
        const double cxp = fma(l, ex, S.pmx[tid]), cyp = fma(l, ey, S.pmy[tid]);  // pt_curr
        // updateDepthFilter ref:482-567
        D3 f_curr{(cxp - P.cx) / P.fx, (cyp - P.cy) / P.fy, 1.0};
        normalize3(f_curr);
        const D3 t{P.ti[0], P.ti[1], P.ti[2]};
        const D3 f2 = qrot(P.qi, f_curr);
        const double b0 = dot3(t, f_ref), b1 = dot3(t, f2);
        const double a00 = dot3(f_ref, f_ref), a01 = -dot3(f_ref, f2), a11 = -dot3(f2, f2);
        const double a10 = -a01;
        // 2x2 solve (the reference uses ColPivHouseholderQR; Cramer differs by O(cond*eps))
        const double det = a00 * a11 - a01 * a10;
        const double ans0 = (b0 * a11 - a01 * b1) / det;
        const double ans1 = (a00 * b1 - a10 * b0) / det;
        const D3 pe{(ans0 * f_ref.x + (t.x + ans1 * f2.x)) / 2.0, (ans0 * f_ref.y + (t.y + ans1 * f2.y)) / 2.0,
                    (ans0 * f_ref.z + (t.z + ans1 * f2.z)) / 2.0};
        const double depth_est = sqrt(dot3(pe, pe));
        const double t_norm = sqrt(dot3(t, t));
        const double alpha = acos(dot3(f_ref, t) / t_norm);
        D3 fcp{(cxp + ex - P.cx) / P.fx, (cyp + ey - P.cy) / P.fy, 1.0};
        normalize3(fcp);
        const D3 mt{-t.x, -t.y, -t.z};
        const double beta_prime = acos(dot3(fcp, mt) / t_norm);
        const double gamma = 3.14159265358979323846 - alpha - beta_prime;
        const double p_prime = t_norm * sin(beta_prime) / sin(gamma);
        const double d_cov = P.inverse_depth ? (1.0 / p_prime - 1.0 / depth_est) : (p_prime - depth_est);
        const double d_cov2 = d_cov * d_cov;
        const double mu0 = P.inverse_depth ? 1.0 / mu : mu;
        const double meas = P.inverse_depth ? (c2 * 1.0 / depth_est) : (c2 * depth_est);
        const double denom = c2 + d_cov2 + 1e-10;
        const double mu_fuse = (d_cov2 * mu0 + meas) / denom;
        const double sig_fuse = (c2 * d_cov2) / denom;
        P.depth[(size_t)y * P.state_pitch + x] = P.inverse_depth ? 1.0 / mu_fuse : mu_fuse;  // ref:560-562
        P.cov2[(size_t)y * P.state_pitch + x] = sig_fuse;                                    // ref:564
    }
    if (P.write_flags && in_img) {
        P.flags[(size_t)y * P.flags_pitch + x] = (uint8_t)((active ? 1 : 0) | (accepted ? 2 : 0));
        const unsigned long long key = S.best[tid];
        P.dbg_ncc[(size_t)y * P.flags_pitch + x] = active ? from_ordered_bits((unsigned)(key >> 32)) : 0.0f;
        const unsigned kb = ((unsigned)key == 0xFFFFFFFFu) ? 0xFFFFu : (0xFFFFFFFEu - (unsigned)key);
        P.dbg_n[(size_t)y * P.flags_pitch + x] = active ? ((n << 16) | (int)(kb > 0xFFFEu ? 0xFFFFu : kb)) : 0;
    }

    // counters: warp reduce -> shared -> one global atomic per CTA
    unsigned int a = __popc(__ballot_sync(0xffffffffu, active));
    unsigned int c = __popc(__ballot_sync(0xffffffffu, accepted));
    unsigned int e = my_evals;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) e += __shfl_down_sync(0xffffffffu, e, o);
    if (lane == 0) {
        atomicAdd(&S.cnt_active, a);
        atomicAdd(&S.cnt_accept, c);
        atomicAdd(&S.cnt_eval, e);
    }
    __syncthreads();
    if (tid == 0) {
        if (S.cnt_active) atomicAdd(&P.counters[0], (unsigned long long)S.cnt_active);
        if (S.cnt_eval) atomicAdd(&P.counters[1], (unsigned long long)S.cnt_eval);
        if (S.cnt_accept) atomicAdd(&P.counters[2], (unsigned long long)S.cnt_accept);
    }
}

// ----------------------------------------------------------------------------------------
// small utility kernels
__global__ void fill_state_kernel(double *depth, double *cov2, size_t n, double d0, double c0) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) { depth[i] = d0; cov2[i] = c0; }
}

// evaludateDepth ref:569-590 over the band: sum of squared errors + count where var < max_variance
__global__ void __launch_bounds__(256) evaluate_depth_kernel(const double *__restrict__ truth, const double *__restrict__ est,
                                                             const double *__restrict__ var, int pitch, int x0, int x1,
                                                             int y0, int y1, double max_variance, double *sum_sq,
                                                             unsigned long long *count) {
    double s = 0;
    unsigned long long n = 0;
    const int w = x1 - x0;
    const long long total = (long long)w * (y1 - y0);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        int y = y0 + (int)(i / w), x = x0 + (int)(i % w);
        double v = var[(size_t)y * pitch + x];
        if (v >= max_variance) continue;  // ref:579 (NaN is counted, as in the reference)
        double e = truth[(size_t)y * pitch + x] - est[(size_t)y * pitch + x];
        s += e * e;
        n++;
    }
    // block reduction
    __shared__ double ss[8];
    __shared__ unsigned long long sn[8];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        s += __shfl_down_sync(0xffffffffu, s, o);
        n += __shfl_down_sync(0xffffffffu, n, o);
    }
    if ((threadIdx.x & 31) == 0) { ss[threadIdx.x >> 5] = s; sn[threadIdx.x >> 5] = n; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double S2 = 0; unsigned long long N = 0;
        for (int i = 0; i < 8; ++i) { S2 += ss[i]; N += sn[i]; }
        atomicAdd(sum_sq, S2);
        atomicAdd(count, N);
    }
}

// getMaskFromVariance ref:199-204: 255 where !(var > max_variance), else 0
__global__ void variance_mask_kernel(const double *__restrict__ var, int pitch, int width, int y0, int y1,
                                     double max_variance, uint8_t *__restrict__ mask, int mask_pitch) {
    int x = blockIdx.x * blockDim.x + threadIdx.x;
    int y = y0 + blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= width || y >= y1) return;
    mask[(size_t)y * mask_pitch + x] = var[(size_t)y * pitch + x] > max_variance ? 0 : 255;
}

}  // namespace dmf
