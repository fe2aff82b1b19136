// Pipe-throughput micro-benchmarks for sm_100a (B200).
//
// Purpose: SURVEY.md §8(d) says the depth-filter path is bound by FP32 issue /
// L1-tex, and MEASURED_PEAKS.json only holds HBM and bf16 numbers.  This file
// measures, on the box, the per-SM issue rates that decide the kernel design
// and that serve as the roofline denominators in bench.py:
//   FFMA, DFMA, IDP4A, IMAD, I2F(u8), PRMT, SHF, LDS, and a few mixes.
// Each kernel runs one full wave (148*k CTAs), every thread executes ITER
// iterations of UNROLL independent dependency chains; rate is reported as
// thread-ops / clk / SM using clock64() deltas of the slowest CTA and also as
// wall-clock ops/s from CUDA events.
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipes pipes.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include <string>
#include <algorithm>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { \
    fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1);} } while (0)

constexpr int ITER = 4096;
constexpr int UNROLL = 8;

struct Out { long long cyc; float sink; };

template <int KIND>
__global__ void __launch_bounds__(512) k_pipe(Out* out, float seedf, int seedi, double seedd) {
    __shared__ float sm[1024];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) sm[i] = seedf * i;
    __syncthreads();
    float f[UNROLL]; int v[UNROLL]; double d[UNROLL];
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) { f[u] = seedf + u + threadIdx.x; v[u] = seedi + u * 77 + threadIdx.x; d[u] = seedd + u; }
    float a = seedf * 1.0001f, b = seedf * 0.5f;
    double da = seedd * 1.0001, db = seedd * 0.5;
    int ia = seedi | 0x01010101, ib = seedi + 3;
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITER; ++it) {
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            if (KIND == 0) {            // FFMA 3-reg
                f[u] = fmaf(f[u], a, b);
            } else if (KIND == 1) {     // DFMA
                d[u] = fma(d[u], da, db);
            } else if (KIND == 2) {     // IDP4A
                v[u] = __dp4a((unsigned)v[u], (unsigned)ia, (unsigned)ib + (unsigned)v[u]);
            } else if (KIND == 3) {     // IMAD
                v[u] = v[u] * ia + ib;
            } else if (KIND == 4) {     // I2F from byte (cvt.rn.f32.u8-like)
                f[u] += (float)((unsigned)__float_as_int(f[u]) & 0xffu);
            } else if (KIND == 5) {     // PRMT
                v[u] = __byte_perm(v[u], ia, 0x4321 + u);
            } else if (KIND == 6) {     // SHF (funnel shift)
                v[u] = __funnelshift_r(v[u], ia, (v[u] & 24));
            } else if (KIND == 7) {     // LDS.32 distinct banks
                f[u] += sm[(threadIdx.x + u * 32 + it) & 1023];
            } else if (KIND == 8) {     // FFMA + IDP4A mix 1:1 (do they share a pipe?)
                f[u] = fmaf(f[u], a, b);
                v[u] = __dp4a((unsigned)v[u], (unsigned)ia, (unsigned)ib + (unsigned)v[u]);
            } else if (KIND == 9) {     // FFMA + DFMA mix 2:1
                f[u] = fmaf(f[u], a, b);
                if ((u & 1) == 0) d[u] = fma(d[u], da, db);
            } else if (KIND == 10) {    // FFMA + PRMT mix 2:1
                f[u] = fmaf(f[u], a, b);
                if ((u & 1) == 0) v[u] = __byte_perm(v[u], ia, 0x4321 + u);
            } else if (KIND == 11) {    // magic-number byte->float: PRMT + FADD
                unsigned m = __byte_perm((unsigned)v[u], 0x4B000000u, 0x7440 + (u & 3));
                f[u] += __int_as_float(m) - 8388608.0f;
                v[u] += it;
            } else if (KIND == 12) {    // IDP4A with independent accumulate (acc chain only)
                v[u] = __dp4a((unsigned)ia + u, (unsigned)ib, (unsigned)v[u]);
            } else if (KIND == 13) {    // IADD3
                v[u] = v[u] + ia + ib;
            } else if (KIND == 14) {    // FFMA + IADD mix 1:1
                f[u] = fmaf(f[u], a, b);
                v[u] = (v[u] ^ ia) + ib;
            } else if (KIND == 15) {    // IDP4A + IADD/LOP mix 1:1
                v[u] = __dp4a((unsigned)v[u], (unsigned)ia, (unsigned)ib);
                v[(u + 1) % UNROLL] ^= ia >> (u + 1);
            } else if (KIND == 16) {    // I2F.F64.S32
                d[u] += (double)(v[u]); v[u] += 1;
            } else if (KIND == 17) {    // F2F f64->f32 + f32->f64
                f[u] = (float)d[u]; d[u] = (double)f[u] + da;
            } else if (KIND == 18) {    // MUFU.RSQ
                f[u] = rsqrtf(f[u]) + a;
            } else if (KIND == 19) {    // LDS.128 broadcast
                float4 q = *reinterpret_cast<const float4*>(&sm[((it + u) * 4) & 1020]);
                f[u] += q.x + q.y + q.z + q.w;
            }
        }
    }
    long long t1 = clock64();
    float s = 0.f;
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) s += f[u] + (float)v[u] + (float)d[u];
    if (threadIdx.x == 0) { out[blockIdx.x].cyc = t1 - t0; }
    if (s == 1234.5678f) out[blockIdx.x].sink = s;
}

struct Case { int kind; const char* name; double ops_per_iter; };

template <int KIND>
static void run(const Case& c, Out* d_out, int blocks, int threads, FILE* js, bool first) {
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    k_pipe<KIND><<<blocks, threads>>>(d_out, 1.000001f, 0x12345678, 1.0000001);
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0));
    k_pipe<KIND><<<blocks, threads>>>(d_out, 1.000001f, 0x12345678, 1.0000001);
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    std::vector<Out> h(blocks);
    CK(cudaMemcpy(h.data(), d_out, sizeof(Out) * blocks, cudaMemcpyDeviceToHost));
    long long mx = 0; for (auto& o : h) mx = std::max(mx, o.cyc);
    int dev; CK(cudaGetDevice(&dev)); cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, dev));
    double ctas_per_sm = (double)blocks / p.multiProcessorCount;
    double ops_per_thread = (double)ITER * UNROLL * c.ops_per_iter;
    double per_clk_sm = ops_per_thread * threads * ctas_per_sm / (double)mx;
    double gops = ops_per_thread * threads * blocks / (ms * 1e-3) * 1e-9;
    double mhz = (double)mx / (ms * 1e-3) * 1e-6;
    printf("%-28s  %8.2f ops/clk/SM   %10.1f Gops/s   (%.3f ms, %lld cyc, ~%.0f MHz)\n",
           c.name, per_clk_sm, gops, ms, mx, mhz);
    fprintf(js, "%s\n  {\"name\": \"%s\", \"ops_per_clk_sm\": %.3f, \"gops\": %.2f, \"ms\": %.4f, \"cycles\": %lld, \"eff_mhz\": %.1f}",
            first ? "" : ",", c.name, per_clk_sm, gops, ms, mx, mhz);
}

int main(int argc, char** argv) {
    const char* path = argc > 1 ? argv[1] : "pipes.json";
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    printf("device: %s, %d SMs, clock %d kHz\n", p.name, p.multiProcessorCount, p.clockRate);
    int blocks = p.multiProcessorCount * 2, threads = 512;
    Out* d_out; CK(cudaMalloc(&d_out, sizeof(Out) * blocks));
    FILE* js = fopen(path, "w");
    fprintf(js, "{\"device\": \"%s\", \"sms\": %d, \"cases\": [", p.name, p.multiProcessorCount);
    Case cs[] = {
        {0, "ffma", 1}, {1, "dfma", 1}, {2, "idp4a_dep", 1}, {3, "imad", 1}, {4, "i2f_u8+fadd+lop", 1},
        {5, "prmt", 1}, {6, "shf+lop", 1}, {7, "lds32", 1}, {8, "ffma+idp4a(pair)", 2}, {9, "ffma+0.5dfma", 1.5},
        {10, "ffma+0.5prmt", 1.5}, {11, "magic_u8_to_f32(prmt+fadd)", 1}, {12, "idp4a_acc", 1}, {13, "iadd3", 1},
        {14, "ffma+lop/iadd(pair)", 2}, {15, "idp4a+lop(pair)", 2}, {16, "i2f_f64", 1}, {17, "f2f_64_32_roundtrip", 1},
        {18, "mufu_rsq+fadd", 1}, {19, "lds128_bcast", 1},
    };
    run<0>(cs[0], d_out, blocks, threads, js, true);
    run<1>(cs[1], d_out, blocks, threads, js, false);
    run<2>(cs[2], d_out, blocks, threads, js, false);
    run<3>(cs[3], d_out, blocks, threads, js, false);
    run<4>(cs[4], d_out, blocks, threads, js, false);
    run<5>(cs[5], d_out, blocks, threads, js, false);
    run<6>(cs[6], d_out, blocks, threads, js, false);
    run<7>(cs[7], d_out, blocks, threads, js, false);
    run<8>(cs[8], d_out, blocks, threads, js, false);
    run<9>(cs[9], d_out, blocks, threads, js, false);
    run<10>(cs[10], d_out, blocks, threads, js, false);
    run<11>(cs[11], d_out, blocks, threads, js, false);
    run<12>(cs[12], d_out, blocks, threads, js, false);
    run<13>(cs[13], d_out, blocks, threads, js, false);
    run<14>(cs[14], d_out, blocks, threads, js, false);
    run<15>(cs[15], d_out, blocks, threads, js, false);
    run<16>(cs[16], d_out, blocks, threads, js, false);
    run<17>(cs[17], d_out, blocks, threads, js, false);
    run<18>(cs[18], d_out, blocks, threads, js, false);
    run<19>(cs[19], d_out, blocks, threads, js, false);
    fprintf(js, "\n]}\n");
    fclose(js);
    CK(cudaFree(d_out));
    return 0;
}
