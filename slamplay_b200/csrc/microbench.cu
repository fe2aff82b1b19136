// microbench.cu — measured per-SM issue rates of the pipes the depth-filter kernels run on (sm_100a).
//
// the path is bound by instruction issue and L1/TEX (LSU) throughput, not by HBM or tensor cores, and
// MEASURED_PEAKS.json holds only HBM and bf16 figures.  bench.py calls dmf_pipe_peaks() in the same process as the timed
// run (clocks logged there) and uses the numbers below as the roofline denominators of ncc_kernel:
//   IDP.4A  (fmaheavy pipe: the u8 x u8 -> s32 cross sums)          thread-ops / clk / SM
//   DFMA    (FP64 pipe: positions, NCC combination)
//   I2F.F64 (XU pipe: int -> double conversions, rsqrt seed)
//   FFMA    (FMA pipe: the "600 FP32 flop per NCC" secondary figure of SURVEY.md §8d)
//   LDG.64 / LDG.128 hitting L1 (LSU data pipe: 128-byte wavefronts / clk / SM)
// Every kernel runs 2 CTAs x 512 threads per SM; a thread executes ITER iterations of UNROLL independent chains.
// Rates come from CUDA events around the launch (ops / s) and from clock64() of the slowest CTA (ops / clk / SM).
#include "../../include/dmf.h"

#include <cuda_runtime.h>
#include <stdint.h>

#include <algorithm>
#include <vector>

namespace {

constexpr int ITER = 2048;
constexpr int UNROLL = 8;
constexpr int THREADS = 512;
constexpr int CTAS_PER_SM = 2;
constexpr int L1_WORDS = 2048;  // 16 KB of uint2 per CTA: stays in L1

struct Out { long long cyc; float sink; };

enum Kind { K_FFMA = 0, K_DFMA, K_IDP4A, K_I2F64, K_LDG64, K_LDG128, K_COUNT };

template <int KIND>
__global__ void __launch_bounds__(THREADS) pipe_kernel(Out *out, const uint2 *__restrict__ buf, float seedf, int seedi, double seedd) {
    float f[UNROLL]; int v[UNROLL]; double d[UNROLL];
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) { f[u] = seedf + u + threadIdx.x; v[u] = seedi + u * 77 + threadIdx.x; d[u] = seedd + u; }
    const float a = seedf * 1.0001f, b = seedf * 0.5f;
    const double da = seedd * 1.0001, db = seedd * 0.5;
    const int ia = seedi | 0x01010101;
    const uint2 *mine = buf + (size_t)blockIdx.x * L1_WORDS;
    const int lane_off = threadIdx.x & 31;
    const int warp_off = (threadIdx.x >> 5) * 64;
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITER; ++it) {
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            if (KIND == K_FFMA) f[u] = fmaf(f[u], a, b);
            else if (KIND == K_DFMA) d[u] = fma(d[u], da, db);
            else if (KIND == K_IDP4A) v[u] = __dp4a((unsigned)ia + u, (unsigned)seedi, (unsigned)v[u]);
            else if (KIND == K_I2F64) { d[u] += (double)(v[u]); v[u] += 1; }
            else if (KIND == K_LDG64) {  // coalesced: the 32 lanes of a warp read 256 contiguous bytes = 2 wavefronts
                const uint2 q = __ldg(mine + ((warp_off + lane_off + (it * UNROLL + u) * 32) & (L1_WORDS - 1)));
                v[u] += (int)(q.x ^ q.y);
            } else if (KIND == K_LDG128) {  // 512 contiguous bytes per warp = 4 wavefronts
                const uint4 q = __ldg(reinterpret_cast<const uint4 *>(mine) + ((warp_off + lane_off + (it * UNROLL + u) * 32) & (L1_WORDS / 2 - 1)));
                v[u] += (int)(q.x ^ q.y ^ q.z ^ q.w);
            }
        }
    }
    const long long t1 = clock64();
    float s = 0.f;
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) s += f[u] + (float)v[u] + (float)d[u];
    if (threadIdx.x == 0) out[blockIdx.x].cyc = t1 - t0;
    if (s == 1234.5678f) out[blockIdx.x].sink = s;
}

template <int KIND>
cudaError_t run_one(Out *d_out, const uint2 *d_buf, int blocks, int n_sm, double ops_per_chain_step, dmf_pipe_rate *r) {
    cudaEvent_t e0, e1;
    cudaError_t e;
    if ((e = cudaEventCreate(&e0)) != cudaSuccess) return e;
    if ((e = cudaEventCreate(&e1)) != cudaSuccess) return e;
    double best_ms = 1e30;
    long long best_cyc = 0;
    std::vector<Out> h(blocks);
    for (int rep = 0; rep < 4; ++rep) {  // rep 0 = warm-up
        cudaEventRecord(e0);
        pipe_kernel<KIND><<<blocks, THREADS>>>(d_out, d_buf, 1.000001f, 0x12345678, 1.0000001);
        cudaEventRecord(e1);
        if ((e = cudaEventSynchronize(e1)) != cudaSuccess) return e;
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        if ((e = cudaMemcpy(h.data(), d_out, sizeof(Out) * blocks, cudaMemcpyDeviceToHost)) != cudaSuccess) return e;
        long long mx = 0;
        for (auto &o : h) mx = std::max(mx, o.cyc);
        if (rep > 0 && ms < best_ms) { best_ms = ms; best_cyc = mx; }
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    const double ops_thread = (double)ITER * UNROLL * ops_per_chain_step;
    r->per_clk_sm = ops_thread * THREADS * ((double)blocks / n_sm) / (double)best_cyc;
    r->per_second = ops_thread * THREADS * blocks / (best_ms * 1e-3);
    r->eff_mhz = (double)best_cyc / (best_ms * 1e-3) * 1e-6;
    return cudaGetLastError();
}

}  // namespace

extern "C" int dmf_pipe_peaks(int device, dmf_pipe_peaks_t *out) {
    if (!out) return DMF_ERR_INVALID;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) return DMF_ERR_CUDA;
    cudaDeviceProp prop{};
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess || cudaSetDevice(device) != cudaSuccess) return DMF_ERR_CUDA;
    const int n_sm = prop.multiProcessorCount, blocks = n_sm * CTAS_PER_SM;
    Out *d_out = nullptr;
    uint2 *d_buf = nullptr;
    if (cudaMalloc(&d_out, sizeof(Out) * blocks) != cudaSuccess) return DMF_ERR_CUDA;
    if (cudaMalloc(&d_buf, sizeof(uint2) * (size_t)L1_WORDS * blocks) != cudaSuccess) { cudaFree(d_out); return DMF_ERR_CUDA; }
    cudaMemset(d_buf, 0x5a, sizeof(uint2) * (size_t)L1_WORDS * blocks);
    cudaError_t e = cudaSuccess;
    out->n_sm = n_sm;
    if (e == cudaSuccess) e = run_one<K_FFMA>(d_out, d_buf, blocks, n_sm, 1.0, &out->ffma);
    if (e == cudaSuccess) e = run_one<K_DFMA>(d_out, d_buf, blocks, n_sm, 1.0, &out->dfma);
    if (e == cudaSuccess) e = run_one<K_IDP4A>(d_out, d_buf, blocks, n_sm, 1.0, &out->idp4a);
    if (e == cudaSuccess) e = run_one<K_I2F64>(d_out, d_buf, blocks, n_sm, 1.0, &out->i2f_f64);
    // LSU: one warp-wide LDG.64 = 2 wavefronts of 128 B, one LDG.128 = 4; reported per THREAD op like the others
    // (per_clk_sm / 32 * wavefronts = wavefronts / clk / SM)
    if (e == cudaSuccess) e = run_one<K_LDG64>(d_out, d_buf, blocks, n_sm, 1.0, &out->ldg64_l1);
    if (e == cudaSuccess) e = run_one<K_LDG128>(d_out, d_buf, blocks, n_sm, 1.0, &out->ldg128_l1);
    cudaFree(d_out);
    cudaFree(d_buf);
    return e == cudaSuccess ? DMF_OK : DMF_ERR_CUDA;
}
