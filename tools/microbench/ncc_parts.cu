// ncc_parts.cu — additive cost model of the ncc_kernel sample loop on sm_100a.
//
// The production kernel (slamplay_b200/csrc/dmf_kernels.cuh: ncc_kernel) spends ~600 SM cycles per
// warp-sample and is insensitive to most single changes, so this micro-benchmark isolates its parts on a
// synthetic but similarly shaped access pattern (adjacent lanes = adjacent pixels, samples 0.7 px apart
// along a near-horizontal line, 8 samples per thread-unit) and times every combination of
//   POS    FP64 position chain + floor / fraction (F2I / I2F on the XU pipe)
//   RAW    24 aligned 32-bit gathers of the 8x8 block + funnel shifts
//   TAB    5 vector loads of the moment table
//   DP4A   56 IDP.4A cross sums + integer centring
//   COMB   FP64 combination (14 I2F.F64, ~45 FP64 ops, rsqrt)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ncc_parts ncc_parts.cu
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "CUDA %s @%d\n", cudaGetErrorString(e_), __LINE__); exit(1);} } while (0)

constexpr int W = 1920, H = 1080, PITCH = 1920;
constexpr int SAMPLES = 8;

__device__ __forceinline__ int dp4(uint32_t a, uint32_t b, int c) { return (int)__dp4a(a, b, (unsigned)c); }

struct Params {
    const uint8_t *img;
    const uint2 *imgx;   // expanded image: imgx[y*W + x] = the 8 bytes img[y][x..x+7]
    const int4 *mom1;
    const int2 *mom2;
    const double4 *units;  // per thread: x0, y0, dx, dy
    double *out;
    int n_units;
};

template <bool POS, int RAW, bool TAB, bool DP4A, bool COMB>
__global__ void __launch_bounds__(256, 2) k_parts(const Params P) {
    const int lane = threadIdx.x & 31;
    double acc = 0;
    uint32_t R0lo[7], R0hi[7], R1lo[7], R1hi[7];
#pragma unroll
    for (int j = 0; j < 7; ++j) {
        R0lo[j] = 0x01020304u * (j + 1) + threadIdx.x; R0hi[j] = (0x00030201u * (j + 2)) & 0x00FFFFFFu;
        R1lo[j] = R0lo[j] << 8; R1hi[j] = __funnelshift_l(R0lo[j], R0hi[j], 8);
    }
    for (int u = blockIdx.x * blockDim.x + threadIdx.x; u < P.n_units; u += gridDim.x * blockDim.x) {
        const double4 un = P.units[u];
        double best = -1.0;
#pragma unroll 1
        for (int k = 0; k < SAMPLES; ++k) {
            double cx, cy;
            int ix, iy;
            double fx, fy;
            if (POS) {
                const double l = fma(0.7, (double)k, -2.45);
                cx = fma(l, un.z, un.x); cy = fma(l, un.w, un.y);
                const bool ok = cx >= 20.0 && cy >= 20.0 && cx + 20.0 < W && cy + 20.0 <= H;
                if (!ok) continue;
                ix = (int)cx; iy = (int)cy;
                fx = cx - (double)ix; fy = cy - (double)iy;
            } else {
                ix = (int)un.x + k; iy = (int)un.y; fx = 0.25 + 0.01 * k; fy = 0.5;
            }
            uint32_t lo[8], hi[8];
            const unsigned off = (unsigned)(iy - 3) * PITCH + (unsigned)(ix - 3);
            if (RAW == 2) {  // expanded image: one aligned 8-byte load per row, no shifts
#pragma unroll
                for (int j = 0; j < 8; ++j) { const uint2 q = __ldg(P.imgx + (size_t)(iy - 3 + j) * W + (ix - 3)); lo[j] = q.x; hi[j] = q.y; }
            } else if (RAW == 1) {
                const unsigned sh = (off & 3u) * 8u;
                const uint32_t *wp = reinterpret_cast<const uint32_t *>(P.img + (off & ~3u));
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const uint32_t w0 = __ldg(wp + j * (PITCH / 4)), w1 = __ldg(wp + j * (PITCH / 4) + 1), w2 = __ldg(wp + j * (PITCH / 4) + 2);
                    lo[j] = __funnelshift_r(w0, w1, sh); hi[j] = __funnelshift_r(w1, w2, sh);
                }
            } else {
#pragma unroll
                for (int j = 0; j < 8; ++j) { lo[j] = off * 2654435761u + j; hi[j] = off * 40503u + j * 7; }
            }
            int4 m00, m10, m01, m11; int2 md;
            if (TAB) {
                const size_t mo = (size_t)(iy - 3) * W + (ix - 3);
                m00 = __ldg(P.mom1 + mo); m10 = __ldg(P.mom1 + mo + 1);
                m01 = __ldg(P.mom1 + mo + W); m11 = __ldg(P.mom1 + mo + W + 1);
                md = __ldg(P.mom2 + mo);
            } else {
                m00 = make_int4(ix, iy, ix + iy, ix - iy); m10 = make_int4(iy, ix + 1, 3 * ix, ix); m01 = make_int4(ix + 2, iy + 2, iy, ix);
                m11 = make_int4(ix, 5 * iy, ix, iy); md = make_int2(ix * 3, iy * 5);
            }
            int cR00, cR10, cR01, cR11;
            if (DP4A) {
                int R00 = 0, R10 = 0, R01 = 0, R11 = 0;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    if (j > 0) { R01 = dp4(R0lo[j - 1], lo[j], dp4(R0hi[j - 1], hi[j], R01)); R11 = dp4(R1lo[j - 1], lo[j], dp4(R1hi[j - 1], hi[j], R11)); }
                    if (j < 7) { R00 = dp4(R0lo[j], lo[j], dp4(R0hi[j], hi[j], R00)); R10 = dp4(R1lo[j], lo[j], dp4(R1hi[j], hi[j], R10)); }
                }
                cR00 = 49 * R00 - 77 * m00.x; cR10 = 49 * R10 - 77 * m10.x; cR01 = 49 * R01 - 77 * m01.x; cR11 = 49 * R11 - 77 * m11.x;
            } else {
                cR00 = (int)(lo[0] ^ hi[7]) + m00.x; cR10 = (int)(lo[1] ^ hi[6]) + m10.x; cR01 = (int)(lo[2] ^ hi[5]) + m01.x; cR11 = (int)(lo[3] ^ hi[4]) + m11.x;
                cR00 += (int)(lo[4] + lo[5] + lo[6] + lo[7]); cR10 += (int)(hi[0] + hi[1] + hi[2] + hi[3]);
            }
            double v;
            if (COMB) {
                const double gx = 1.0 - fx, gy = 1.0 - fy;
                const double w00 = gx * gy, w10 = fx * gy, w01 = gx * fy, w11 = fx * fy;
                double num = w00 * (double)cR00; num = fma(w10, (double)cR10, num); num = fma(w01, (double)cR01, num); num = fma(w11, (double)cR11, num);
                const double g0000 = (double)m00.y, g1010 = (double)m10.y, g0101 = (double)m01.y, g1111 = (double)m11.y;
                const double g0010 = (double)m00.z, g0111 = (double)m01.z, g0001 = (double)m00.w, g1011 = (double)m10.w;
                const double g0011 = (double)md.x, g1001 = (double)md.y;
                double a0 = w00 * g0000; a0 = fma(w10, g0010, a0); a0 = fma(w01, g0001, a0); a0 = fma(w11, g0011, a0);
                double a1 = w00 * g0010; a1 = fma(w10, g1010, a1); a1 = fma(w01, g1001, a1); a1 = fma(w11, g1011, a1);
                double a2 = w00 * g0001; a2 = fma(w10, g1001, a2); a2 = fma(w01, g0101, a2); a2 = fma(w11, g0111, a2);
                double a3 = w00 * g0011; a3 = fma(w10, g1011, a3); a3 = fma(w01, g0111, a3); a3 = fma(w11, g1111, a3);
                double den2 = w00 * a0; den2 = fma(w10, a1, den2); den2 = fma(w01, a2, den2); den2 = fma(w11, a3, den2);
                const double dd = fma(12345.0, fabs(den2), 1015.2);
                v = num * rsqrt(dd);
            } else {
                v = (double)(cR00 + cR10 + cR01 + cR11 + m00.y + m10.y + m01.y + m11.y + m00.z + m01.z + m00.w + m10.w + md.x + md.y) * fx;
            }
            if (v > best) best = v;
        }
        acc += best;
    }
    if (acc == 1.2345e-300) P.out[lane] = acc;
}

template <bool POS, int RAW, bool TAB, bool DP4A, bool COMB>
static void run(const char *name, const Params &P, int grid) {
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    k_parts<POS, RAW, TAB, DP4A, COMB><<<grid, 256>>>(P);
    CK(cudaDeviceSynchronize());
    float best = 1e9f;
    for (int r = 0; r < 3; ++r) {
        CK(cudaEventRecord(e0));
        k_parts<POS, RAW, TAB, DP4A, COMB><<<grid, 256>>>(P);
        CK(cudaEventRecord(e1)); CK(cudaDeviceSynchronize());
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); best = ms < best ? ms : best;
    }
    const double samples = (double)P.n_units * SAMPLES;
    // cycles per warp-sample per scheduler at 1.9 GHz: time * f * (148*4 schedulers) / (samples/32)
    const double cyc = best * 1e-3 * 1.9e9 * 148 * 4 / (samples / 32);
    printf("%-34s %8.3f ms   %6.1f G samples/s   %6.0f cycles per warp-sample per scheduler\n", name, best, samples / best * 1e-6, cyc);
}

int main() {
    std::vector<uint8_t> img((size_t)PITCH * H);
    srand(1);
    for (auto &b : img) b = (uint8_t)(rand() & 255);
    uint8_t *d_img; CK(cudaMalloc(&d_img, img.size())); CK(cudaMemcpy(d_img, img.data(), img.size(), cudaMemcpyHostToDevice));
    std::vector<uint2> imgx((size_t)W * H);
    for (int y = 0; y < H; ++y) for (int x = 0; x + 8 <= W; ++x) { uint2 q; memcpy(&q, &img[(size_t)y * PITCH + x], 8); imgx[(size_t)y * W + x] = q; }
    uint2 *d_imgx; CK(cudaMalloc(&d_imgx, imgx.size() * 8)); CK(cudaMemcpy(d_imgx, imgx.data(), imgx.size() * 8, cudaMemcpyHostToDevice));
    int4 *d_m1; int2 *d_m2;
    CK(cudaMalloc(&d_m1, (size_t)W * H * sizeof(int4))); CK(cudaMalloc(&d_m2, (size_t)W * H * sizeof(int2)));
    CK(cudaMemset(d_m1, 1, (size_t)W * H * sizeof(int4))); CK(cudaMemset(d_m2, 1, (size_t)W * H * sizeof(int2)));
    // units: one per interior pixel, in row-major order with a small random parallax, like a steady-state frame
    const int wi = W - 140, hi = H - 60;
    const int n_units = wi * hi;
    std::vector<double> units((size_t)n_units * 4);
    for (int y = 0; y < hi; ++y)
        for (int x = 0; x < wi; ++x) {
            const size_t i = ((size_t)y * wi + x) * 4;
            const double par = (rand() % 2000) / 100.0 - 10.0;  // +-10 px parallax scatter between neighbours
            units[i] = 70.0 + x + par; units[i + 1] = 30.0 + y + (rand() % 100) / 50.0;
            units[i + 2] = 0.9995; units[i + 3] = 0.0316;
        }
    double *d_units; CK(cudaMalloc(&d_units, units.size() * 8)); CK(cudaMemcpy(d_units, units.data(), units.size() * 8, cudaMemcpyHostToDevice));
    double *d_out; CK(cudaMalloc(&d_out, 64 * 8));
    Params P{d_img, d_imgx, d_m1, d_m2, reinterpret_cast<const double4 *>(d_units), d_out, n_units};
    int grid = 148 * 2;
    printf("units %d, samples %d\n", n_units, n_units * SAMPLES);
    run<true, 1, true, true, true>("all", P, grid);
    run<true, 2, true, true, true>("all, expanded raw (LDG.64)", P, grid);
    run<true, 0, true, true, true>("all - RAW", P, grid);
    run<true, 1, false, true, true>("all - TAB", P, grid);
    run<true, 2, false, true, true>("all - TAB, expanded raw", P, grid);
    run<true, 0, false, true, true>("all - RAW - TAB", P, grid);
    run<false, 1, false, false, false>("RAW only", P, grid);
    run<false, 2, false, false, false>("expanded RAW only", P, grid);
    run<false, 0, true, false, false>("TAB only", P, grid);
    run<false, 1, true, false, false>("RAW + TAB", P, grid);
    run<false, 2, true, false, false>("expanded RAW + TAB", P, grid);
    return 0;
}
