// dmf_kernels.cuh — sm_100a kernels of the dense monocular depth filter.
//
// Replaces, for every interior pixel of the reference frame and each new frame, the CPU
// loop of luigifreda/slamplay dense_mapping/test_monocular_mapping.cpp ("ref:LINE"):
//   update ref:355-393, epipolarSearch ref:397-447, NCC ref:449-480,
//   getBilinearInterpolatedValue ref:165-174, updateDepthFilter ref:482-567.
//
// Once per reference frame: ref_stats_kernel (patch sums) and ref_expand_kernel (7-byte rows as aligned 64-bit words).
// Per update three kernels run; the fusion of update k is deferred into the set-up of update k+1:
//
//   advance_kernel (thread = slot of the previous update, FP64)  accept test ref:443, triangulation, uncertainty and
//                 Gaussian fusion ref:482-567 of update k-1, in place on the HBM-resident maps; then, straight from
//                 the fused registers, the set-up of update k: gate ref:366, projections of mu and mu±3σ ref:402-422,
//                 trip count n of the l-loop ref:432.  The n samples of a pixel are cut into work UNITS of at most
//                 CHUNK consecutive samples; units are appended to per-length lists in HBM (length CHUNK first, ...,
//                 length 1 last).  Active pixels live in SLOTS: every 32x8 tile keeps its active pixels packed at the
//                 front of its own slot range.  setup_kernel (thread = pixel, state from the maps) and fuse_kernel
//                 (thread = slot) are the two halves as stand-alone kernels: first update after the maps were
//                 loaded, strict drop-in mode, debug planes, and whenever the maps are read (flush).
//   moments_kernel (thread = column, sliding 7-row window)  the frame-only integer moments of every 8x8 block
//                 position (see "NCC arithmetic"); as a by-product the "expanded" frame: the 8 bytes [x, x+8) of
//                 every row position as one aligned 64-bit word, so that a sample fetches each row of its 8x8 block
//                 with one LDG.64.  Runs on its own stream beside the previous update (double-buffered tables).
//   ncc_kernel     (thread = unit)  persistent CTAs pull 64-unit grabs of the lists with an
//                 atomic cursor, so every warp runs units of ONE length (no divergence on the
//                 search length, which varies 0..286 per pixel) and the chip stays balanced
//                 whatever the spatial distribution of converged / diverged pixels.  The unit's
//                 7x7 reference patch lives in registers for all its samples; its best sample
//                 goes to the slot's 64-bit arg-max key with one atomicMax.
//
// NCC arithmetic.  All 49 taps of one NCC share the same four bilinear weights (the tap
// offsets are integers, ref:461), so every sum the ZNCC needs is a linear / quadratic form
// in (w00,w10,w01,w11) over INTEGER moments of the 8x8 u8 block under the sample: 4 window
// sums S, 4 cross sums R with the reference patch, 10 Gram sums G, all exact in int32, centred
// exactly in int32 (49*R - Sr*S, 49*G - S*S').
//   * S and the 10 centred Gram sums depend only on the current frame and the integer position
//     of the block: moments_kernel computes them ONCE per frame for every block position
//     (IDP.4A, 4 u8 MACs per instruction) into a 20 B/px table in HBM; a sample reads 5 vector
//     loads from it instead of redoing ~136 dp4a.
//   * the 4 cross sums need the reference patch: 56 IDP.4A per sample on manual __ldg gathers of
//     the block (texture units filter with 8-bit weights and would break parity).
//   * the final combination (bilinear weights, w^T G w in separable form, 1/sqrt) runs in FP64 on the otherwise
//     idle FP64 pipe, so the NCC agrees with the reference's two-pass FP64 ZNCC to ~1e-13 and
//     the arg-max / 0.85 decisions are the reference's except for exact ties.
// Adjacent lanes hold adjacent pixels at the same chunk, so their gathers fall into the same
// cache lines whatever the direction of the epipolar line.
// The arg-max keeps the reference's "first strict maximum" (ref:438): 64-bit key =
// (52-bit mantissa of ncc + 3.0, i.e. fixed point with 2^-51 resolution) << 9 | (510 - sample
// index); low bits 511 mark the "no winner yet" sentinel that goes with best_ncc = -1.0 (ref:430).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "dmf_geometry.h"

namespace dmf {

constexpr int TILE_W = 32;
#ifndef DMF_TILE_H
#define DMF_TILE_H 8
#endif
constexpr int TILE_H = DMF_TILE_H;
constexpr int TILE_PIX = TILE_W * TILE_H;
#ifndef DMF_CHUNK
#define DMF_CHUNK 16
#endif
#ifndef DMF_GRAB
#define DMF_GRAB 64
#endif
constexpr int CHUNK = DMF_CHUNK;  // samples per work unit
constexpr int CHUNK_BITS = 6;     // unit = (pixel index << CHUNK_BITS) | chunk index   (chunk < 64)
#ifndef DMF_NCC_THREADS
#define DMF_NCC_THREADS 192
#endif
#ifndef DMF_NCC_MIN_BLOCKS
#define DMF_NCC_MIN_BLOCKS 3
#endif
constexpr int NCC_THREADS = DMF_NCC_THREADS;
constexpr int GRAB = DMF_GRAB;    // units a warp pulls per atomic
constexpr int NCC_AREA = 49;
// 1e-10 * (49*255^2)^2 : the reference's epsilon (ref:479) in centred-integer units
constexpr double NCC_EPS_INT = 1015.2029750625;
constexpr unsigned KEY_IDX_BITS = 9;          // sample index < 510
constexpr unsigned KEY_SENTINEL_LO = 511u;

// Per-frame control block in HBM.
struct Ctrl {
    unsigned int count[CHUNK + 1];  // count[L] = number of units of length L (L = 1..CHUNK)
    unsigned int cursor;            // next unclaimed slot of the padded, concatenated lists
    unsigned int pad[6];
};

// Record of one ACTIVE pixel of one frame (64 B = half a cache line, four 16-byte vectors): everything
// ncc_kernel / fuse_kernel need about it, so a work unit starts with ONE line fetch.  Records are compacted:
// every active pixel gets a slot in its band's slot range (CTA-aggregated counters), units and the fusion address slots.
struct __align__(16) PixelRec {
    double2 pm;        // px_mean_curr ref:406
    double2 dir;       // epipolar_direction ref:419-420
    double half;       // half_length ref:421-422
    int nSr;           // -(sum of the 49 reference bytes)
    int den1;          // 49*sum r^2 - (sum r)^2
    int4 xy;           // pixel coordinates (x, y, -, -)
};
static_assert(sizeof(PixelRec) == 64, "PixelRec must be 64 bytes");

typedef int mom2_t;   // the two diagonal Gram terms enter the NCC with the same weight: the table carries their sum
// Expanded current frame: currx[y*W + x] = bytes curr[y][x..x+7] as one aligned 64-bit word (a sample fetches each row of
// its 8x8 block with one LDG.64).  Storing row PAIRS (16 B, four LDG.128 per sample) was measured: -3 % on ncc_kernel at
// 1080p, +30 % on the replicated moments kernel, no gain at 8 GPUs (profiles/r02_ab_rowpairs.txt, commit f0b8015).
typedef uint2 currx_t;

struct KParams {
    int width, height, border;
    // Rows owned by this context: local row rl in [0, n_rows) is image row
    //   y = row0 + ((rl / blk) * cyc + pos) * blk + rl % blk,   pos = ph (even rl / blk) or cyc-1-ph (odd)
    // (block-cyclic: blocks of `blk` rows dealt to `cyc` contexts in boustrophedon order, so a load that
    //  varies linearly down the image splits evenly; this context is number `ph`;
    //  a contiguous band is blk = n_rows, cyc = 1, ph = 0).
    int row0, blk, cyc, ph, n_rows;
    int rev_round;  // index of the incomplete last round (dealt from the highest context down whatever its parity), or -1
    int wi;                  // width - 2*border
    int n_pix;               // interior pixels owned = wi * n_rows
    int inverse_depth;
    int write_flags;
    double ncc_thresh;
    double fx, fy, cx, cy;
    double step, max_half_len, min_depth, n_sigma, min_cov, max_cov;
    double q[4], t[3];    // T_C_R (unit quaternion x,y,z,w + translation)
    double qi[4], ti[3];  // T_R_C = T_C_R^-1 (ref:491), computed on the host
    double ti_norm;       // |t_RC| (ref:525)
    double bd, wd, hd;    // border, width, height as doubles (inside() ref:222-224 without per-sample I2F)
    const uint8_t *curr;  // pitched, 4-byte aligned rows
    const currx_t *currx; // expanded current frame (see currx_t; written by moments_kernel)
    const uint8_t *ref;
    const uint2 *refx;    // expanded reference frame: refx[y*width + x] = bytes ref[y][x-3 .. x+3], 0   (ref_expand_kernel)
    const int2 *refstat;  // per pixel: (sum r, 49*sum r^2 - (sum r)^2)
    const int4 *mom1;     // per block position of the current frame: {S, cQ, cH, cV}   (moments_kernel)
    const mom2_t *mom2;   //                                          cD1 + cD2
    double *depth;
    double *cov2;
    // per-frame scratch, indexed by the slot of the active pixel (compacted by setup_kernel)
    PixelRec *rec;                                 // written by setup_kernel
    double2 *state_c;                              // (depth, cov2) of the slot's pixel as setup_kernel read them
    unsigned long long *best;                      // arg-max keys
    unsigned int *units_full;                      // units of length CHUNK
    unsigned int *units_tail;                      // (CHUNK-1) lists of capacity n_pix: lengths 1..CHUNK-1
    Ctrl *ctrl;                                    // control block of the frame being set up / searched
    Ctrl *ctrl_zero;                               // control block re-armed by the kernel that finishes a frame
    // Slot space: CTA c of the pixel grid (a TILE_W x TILE_H tile of the band) owns the slots [c * TILE_PIX,
    // (c + 1) * TILE_PIX) for the whole sequence and keeps its cta_cnt[c] active pixels packed at the front, in tile
    // scan order.  No counter is shared between CTAs, and slot order never drifts away from image order (with shared
    // counters the arrival-order shuffle compounds from update to update: measured -28 % in ncc_kernel from the lost
    // L1/L2 locality, -17 % even with one counter per 8-row band).
    unsigned int *cta_cnt;                         // of the update being set up
    // the frame being FINISHED by fuse_kernel / advance_kernel (records, state, keys, per-CTA counts of its slots)
    const PixelRec *rec_fin;
    const double2 *state_fin;
    const unsigned long long *best_fin;
    const unsigned int *cta_fin;
    uint8_t *flags;
    float *dbg_ncc;  // with write_flags: best NCC per active pixel (rounded to f32)
    int *dbg_n;      // with write_flags: (trip count << 16) | winning iteration (0xFFFF: none)
    unsigned long long *counters;  // [0]=active [1]=ncc_evals [2]=accepted
    int curr_pitch, ref_pitch, stat_pitch, state_pitch, flags_pitch, mom_pitch;  // in elements
};

// ----------------------------------------------------------------------------------------
// FP64 geometry: dmf_geometry.h (the reference's operation order, no FMA contraction, IEEE div / sqrt)
typedef dmf_geom::V3 D3;
__device__ __forceinline__ dmf_geom::Camera camera_of(const KParams &P) { return {P.fx, P.fy, P.cx, P.cy}; }
// exact int -> double for |k| < 2^31 without the slow I2F.F64 path
__device__ __forceinline__ double int2double_fast(int k) {
    return __hiloint2double(0x43300000, (int)((unsigned)k ^ 0x80000000u)) - 4503601774854144.0;  // 2^52 + 2^31
}
__device__ __forceinline__ int row_of(const KParams &P, int rl) {
    const int b = rl / P.blk;
    const int pos = ((b & 1) || b == P.rev_round) ? (P.cyc - 1 - P.ph) : P.ph;
    return P.row0 + (b * P.cyc + pos) * P.blk + (rl - b * P.blk);
}
// arg-max key: ncc in [-1,1] -> ncc + 3.0 in [2,4): the 52 mantissa bits are an order-preserving
// fixed-point code with 2^-51 resolution.
__device__ __forceinline__ unsigned long long ncc_key(double ncc, int k) {
    double c = fmin(fmax(ncc, -1.0), 1.0) + 3.0;
    if (c >= 4.0) c = 3.9999999999999996;
    const unsigned long long m = (unsigned long long)__double_as_longlong(c) & 0x000FFFFFFFFFFFFFull;
    return (m << KEY_IDX_BITS) | (unsigned long long)(510u - (unsigned)k);
}
__device__ __forceinline__ double key_ncc(unsigned long long key) {
    return __longlong_as_double((long long)((key >> KEY_IDX_BITS) | 0x4000000000000000ull)) - 3.0;
}
__device__ __forceinline__ bool key_has_winner(unsigned long long key) { return ((unsigned)key & 511u) != KEY_SENTINEL_LO; }
__device__ __forceinline__ int key_index(unsigned long long key) { return 510 - (int)((unsigned)key & 511u); }
__device__ __forceinline__ unsigned long long key_init() { return (unsigned long long)KEY_SENTINEL_LO; }  // best_ncc = -1.0 ref:430

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

// Once per reference frame: the 7 bytes [x-3, x+3] of every row position as one aligned 64-bit word (8th byte 0), so
// that a work unit fetches each row of its 7x7 reference patch with ONE aligned LDG.64 and no funnel shifts.
__global__ void __launch_bounds__(256) ref_expand_kernel(const uint8_t *__restrict__ ref, int ref_pitch, int width, int height,
                                                         uint2 *__restrict__ refx) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y;
    if (x < 3 || x + 3 >= width || y >= height) return;
    const uint8_t *row = ref + (size_t)y * ref_pitch + (x - 3);
    const uint32_t lo = row[0] | (row[1] << 8) | (row[2] << 16) | ((uint32_t)row[3] << 24);
    const uint32_t hi = row[4] | (row[5] << 8) | (row[6] << 16);
    refx[(size_t)y * width + x] = make_uint2(lo, hi);
}

// ----------------------------------------------------------------------------------------
// Per-pixel setup of one update, shared by setup_kernel (thread = pixel, state read from the maps) and advance_kernel
// (thread = slot of the previous update, state straight from its fusion).
struct PixelWork {
    bool active;
    int x, y, n;                         // pixel, trip count of the search loop
    double mu, c2;                       // state (ref:366)
    double pmx, pmy, lx, ly, half;       // epipolar segment
    int2 st;                             // reference-patch statistics
    D3 f_ref;                            // unit ray of the pixel (valid if have_ray)
    bool have_ray;
};
__device__ __forceinline__ void prepare_pixel(const KParams &P, PixelWork &w, bool have) {
    w.n = 0; w.pmx = w.pmy = w.lx = w.ly = w.half = 0; w.st = make_int2(0, 0);
    w.active = have && !(w.c2 < P.min_cov || w.c2 > P.max_cov);  // ref:366 — NaN passes the gate
    if (w.active) {
        const dmf_geom::Camera cam = camera_of(P);
        if (!w.have_ray) w.f_ref = dmf_geom::unit_ray(cam, (double)w.x, (double)w.y);  // ref:402-403
        const dmf_geom::Segment sg = dmf_geom::search_segment(cam, P.q, P.t, w.f_ref, w.mu, sqrt(w.c2) /* ref:377 */, P.n_sigma,
                                                              P.min_depth, P.max_half_len, P.inverse_depth != 0);
        w.pmx = sg.pm.x; w.pmy = sg.pm.y; w.lx = sg.dir.x; w.ly = sg.dir.y; w.half = sg.half;
        w.n = dmf_geom::trip_count(w.half, P.step);  // `for (l = -half; l <= half; l += step)` ref:432, accumulated l
        w.st = __ldg(&P.refstat[(size_t)w.y * P.stat_pitch + w.x]);
    }
}

// Compaction + work-unit emission; must be reached by every thread of the CTA (two barriers).  Space is claimed once
// per CTA and per list (thread L does the atomicAdd for list L, thread 0 the one for the active-pixel slots, so the
// latency is paid once per CTA instead of by every warp); every warp then writes its units chunk-major so that
// adjacent list entries hold neighbouring pixels.
// `accepted` / `n_fin` (advance_kernel): the finished frame's counters ride on the same CTA-wide aggregation.
__device__ __forceinline__ void emit_pixel(const KParams &P, const PixelWork &w, bool accepted = false, unsigned n_fin = 0) {
    __shared__ unsigned s_cnt[TILE_PIX / 32][CHUNK + 2];   // [warp][L]: units of length L (L == CHUNK: full units); [warp][0]: active pixels
    __shared__ unsigned s_base[TILE_PIX / 32][CHUNK + 1];  // first entry of that warp in list L / first slot
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned lt_mask = (1u << lane) - 1u;
    const int n_full = w.n / CHUNK, tail = w.n % CHUNK;
    const int tot_full = __reduce_add_sync(0xffffffffu, n_full);
    const int m_full = __reduce_max_sync(0xffffffffu, n_full);
    // lanes with the same tail length form a group (one MATCH instead of CHUNK-1 ballots): rank inside the group,
    // and the group's first lane posts its size
    const unsigned peers = __match_any_sync(0xffffffffu, tail);
    const unsigned my_rank = (unsigned)__popc(peers & lt_mask);
    const unsigned act_bal = __ballot_sync(0xffffffffu, w.active);
    const unsigned acc_bal = __ballot_sync(0xffffffffu, accepted);
    if (lane <= CHUNK + 1) s_cnt[warp][lane] = (lane == 0) ? (unsigned)__popc(act_bal) : (lane == CHUNK) ? (unsigned)tot_full : (lane == CHUNK + 1) ? (unsigned)__popc(acc_bal) : 0u;
    __syncwarp();
    if (tail > 0 && my_rank == 0) s_cnt[warp][tail] = (unsigned)__popc(peers);
    __syncthreads();
    if (tid <= CHUNK) {
        unsigned pre[TILE_PIX / 32], tot = 0;
#pragma unroll
        for (int k = 0; k < TILE_PIX / 32; ++k) { pre[k] = tot; tot += s_cnt[k][tid]; }
        const unsigned cta = blockIdx.y * gridDim.x + blockIdx.x;
        unsigned base = cta * TILE_PIX;                     // tid 0: the CTA's own slot range
        if (tid == 0) P.cta_cnt[cta] = tot;
        else base = tot ? atomicAdd(&P.ctrl->count[tid], tot) : 0u;
#pragma unroll
        for (int k = 0; k < TILE_PIX / 32; ++k) s_base[k][tid] = base + pre[k];
    } else if (tid == CHUNK + 1 && n_fin) {  // counters of the finished frame: active, accepted
        unsigned acc = 0;
#pragma unroll
        for (int k = 0; k < TILE_PIX / 32; ++k) acc += s_cnt[k][CHUNK + 1];
        atomicAdd(&P.counters[0], (unsigned long long)n_fin);
        if (acc) atomicAdd(&P.counters[2], (unsigned long long)acc);
    }
    __syncthreads();
    const unsigned slot = s_base[warp][0] + (unsigned)__popc(act_bal & lt_mask);
    if (w.active) {
        PixelRec *rec = P.rec + slot;  // four 16-byte vector stores
        rec->pm = make_double2(w.pmx, w.pmy);
        rec->dir = make_double2(w.lx, w.ly);
        *reinterpret_cast<int4 *>(&rec->half) = make_int4(__double2loint(w.half), __double2hiint(w.half), -w.st.x, w.st.y);
        rec->xy = make_int4(w.x, w.y, 0, 0);
        P.state_c[slot] = make_double2(w.mu, w.c2);
        P.best[slot] = key_init();
    }
    if (m_full > 0) {
        unsigned base = s_base[warp][CHUNK];
        for (int j = 0; j < m_full; ++j) {
            const unsigned bal = __ballot_sync(0xffffffffu, n_full > j);
            if (n_full > j) P.units_full[base + __popc(bal & lt_mask)] = (slot << CHUNK_BITS) | (unsigned)j;
            base += __popc(bal);
        }
    }
    if (tail > 0)
        P.units_tail[(size_t)(tail - 1) * P.n_pix + s_base[warp][tail] + my_rank] = (slot << CHUNK_BITS) | (unsigned)n_full;
}

// K2a: setup of an update from the maps (thread = pixel): the first update after the state was loaded, or any update
// whose predecessor has already been fused (strict drop-in mode, debug planes on).
__global__ void __launch_bounds__(TILE_PIX) setup_kernel(const __grid_constant__ KParams P) {
    const int tid = threadIdx.x;
    PixelWork w;
    w.x = P.border + blockIdx.x * TILE_W + (tid % TILE_W);
    const int rl = blockIdx.y * TILE_H + (tid / TILE_W);
    const bool in_img = (w.x < P.width - P.border) && (rl < P.n_rows);
    w.y = row_of(P, rl);
    w.mu = 0; w.c2 = 0; w.have_ray = false;
    if (in_img) {
        w.c2 = P.cov2[(size_t)w.y * P.state_pitch + w.x];
        w.mu = P.depth[(size_t)w.y * P.state_pitch + w.x];  // issued with the cov load: one round trip
    }
    prepare_pixel(P, w, in_img);
    if (in_img && P.write_flags) {  // debug planes: fuse_kernel only visits active pixels
        const size_t o = (size_t)w.y * P.flags_pitch + w.x;
        P.dbg_n[o] = w.n;
        if (!w.active) { P.flags[o] = 0; P.dbg_ncc[o] = 0.0f; }
    }
    emit_pixel(P, w);
}

// ----------------------------------------------------------------------------------------
__device__ __forceinline__ int dp4(uint32_t a, uint32_t b, int c) { return (int)__dp4a(a, b, (unsigned)c); }

// K2m: once per current frame — the frame-only part of every possible NCC: window sum and centred
// Gram sums of the 8x8 block at each position (x,y) = top-left tap.  A sample whose top-left tap is
// (bx,by) needs   mom1 at (bx,by),(bx+1,by),(bx,by+1),(bx+1,by+1)  and  mom2 at (bx,by):
//   mom1(x,y) = { S(x,y), 49*Q - S^2, 49*H - S(x,y)S(x+1,y), 49*V - S(x,y)S(x,y+1) }   (Q,H,V: squares,
//   mom2(x,y) = (49*D1 - S(x,y)S(x+1,y+1)) + (49*D2 - S(x+1,y)S(x,y+1))                 horizontal / vertical /
//                                                                                        diagonal neighbour products)
// A thread owns one column x and slides the 7-row window down a strip of MOM_STRIP rows: per step the
// row (pair) leaving the window is subtracted and the row (pair) entering it is added, 28 IDP.4A per
// position instead of 176 for a from-scratch 7x7 evaluation.  Adjacent threads hold adjacent columns,
// so the row loads and the 20-byte table stores are coalesced.
#ifndef DMF_MOM_STRIP
#define DMF_MOM_STRIP 16
#endif
#ifndef DMF_MOM_THREADS
#define DMF_MOM_THREADS 128
#endif
constexpr int MOM_STRIP = DMF_MOM_STRIP;
constexpr int MOM_THREADS = DMF_MOM_THREADS;

struct RowBytes { uint32_t x0l, x0h, x1l, x1h, hi; };  // columns 0..6 (x0) and 1..7 (x1) of one block row; hi = bytes 4..7
// SMEM: the row lives in shared memory (moments_bulk_kernel) instead of global memory
template <bool SMEM>
__device__ __forceinline__ RowBytes load_row_bytes(const uint32_t *wp, unsigned sh) {
    uint32_t w0, w1, w2;
    if (SMEM) { w0 = wp[0]; w1 = wp[1]; w2 = wp[2]; }
    else { w0 = __ldg(wp); w1 = __ldg(wp + 1); w2 = __ldg(wp + 2); }
    const uint32_t lo = __funnelshift_r(w0, w1, sh), hi = __funnelshift_r(w1, w2, sh);
    return {lo, hi & 0x00FFFFFFu, __funnelshift_r(lo, hi, 8), hi >> 8, hi};
}
struct RowSums { int s0, s1, q, h; };  // one row: sum cols 0..6, sum cols 1..7, sum of squares cols 0..6, neighbour products
__device__ __forceinline__ RowSums row_sums(const RowBytes &r) {
    const uint32_t ONES = 0x01010101u;
    return {dp4(r.x0l, ONES, dp4(r.x0h, ONES, 0)), dp4(r.x1l, ONES, dp4(r.x1h, ONES, 0)),
            dp4(r.x0l, r.x0l, dp4(r.x0h, r.x0h, 0)), dp4(r.x0l, r.x1l, dp4(r.x0h, r.x1h, 0))};
}
struct PairSums { int v, d1, d2; };  // rows (a, a+1): vertical and the two diagonal neighbour products
__device__ __forceinline__ PairSums pair_sums(const RowBytes &a, const RowBytes &b) {
    return {dp4(a.x0l, b.x0l, dp4(a.x0h, b.x0h, 0)), dp4(a.x0l, b.x1l, dp4(a.x0h, b.x1h, 0)),
            dp4(a.x1l, b.x0l, dp4(a.x1h, b.x0h, 0))};
}

// One column x, positions y0 .. y_end-1: wp = the 4-byte aligned word that holds byte (x, y0), pw = words per row,
// sh = 8 * (byte offset of x inside that word).  Rows up to min(y_end + 8, height - 1) are read.
template <bool SMEM>
__device__ __forceinline__ void moments_strip(const uint32_t *wp, int pw, unsigned sh, int x, int y0, int y_end, int width, int height,
                                              int4 *__restrict__ mom1, mom2_t *__restrict__ mom2, int mom_pitch, currx_t *__restrict__ currx) {
    // window state for position y0: single-row sums over rows y0..y0+6, pair sums over (y0,y0+1)..(y0+6,y0+7)
    int S0 = 0, S1 = 0, Q = 0, H = 0, V = 0, D1 = 0, D2 = 0;
    RowBytes prev = load_row_bytes<SMEM>(wp, sh);
#pragma unroll
    for (int j = 0; j < 7; ++j) {
        const RowBytes next = load_row_bytes<SMEM>(wp + (j + 1) * pw, sh);
        const RowSums r = row_sums(prev);
        const PairSums p = pair_sums(prev, next);
        S0 += r.s0; S1 += r.s1; Q += r.q; H += r.h; V += p.v; D1 += p.d1; D2 += p.d2;
        prev = next;
    }
    // sliding: rows leaving (old_a = y, old_b = y+1) and entering (new_a = y+7, new_b = y+8)
    RowBytes old_a = load_row_bytes<SMEM>(wp, sh), old_b = load_row_bytes<SMEM>(wp + pw, sh);
    RowBytes new_a = prev;  // row y0+7
    RowBytes new_b = load_row_bytes<SMEM>(wp + 8 * pw, sh);
    for (int y = y0; y < y_end; ++y) {
        const RowSums ro = row_sums(old_a), rn = row_sums(new_a);
        const PairSums po = pair_sums(old_a, old_b), pn = pair_sums(new_a, new_b);
        const int S0n = S0 - ro.s0 + rn.s0, S1n = S1 - ro.s1 + rn.s1;  // window sums of position y+1: S01, S11 of y
        int4 a;
        a.x = S0;
        a.y = NCC_AREA * Q - S0 * S0;
        a.z = NCC_AREA * H - S0 * S1;
        a.w = NCC_AREA * V - S0 * S0n;
        const int bx = NCC_AREA * D1 - S0 * S1n, by = NCC_AREA * D2 - S1 * S0n;
        mom1[(size_t)y * mom_pitch + x] = a;
        mom2[(size_t)y * mom_pitch + x] = bx + by;  // |bx|, |by| < 2^30: exact
        // by-product, "sliding window expansion": the 8 bytes [x, x+8) of row y as one aligned 64-bit word, so
        // that a sample fetches each row of its 8x8 block with ONE aligned LDG.64 instead of three LDG.32 + two
        // funnel shifts
        currx[(size_t)y * width + x] = make_uint2(old_a.x0l, old_a.hi);
        // advance the window to position y+1
        S0 = S0n; S1 = S1n;
        Q += rn.q - ro.q; H += rn.h - ro.h;
        V += pn.v - po.v; D1 += pn.d1 - po.d1; D2 += pn.d2 - po.d2;
        old_a = old_b; new_a = new_b;
        const int yn = y + 1;
        old_b = load_row_bytes<SMEM>(wp + (size_t)(yn + 1 - y0) * pw, sh);
        new_b = load_row_bytes<SMEM>(wp + (size_t)(min(yn + 8, height - 1) - y0) * pw, sh);
    }
}

__global__ void __launch_bounds__(MOM_THREADS) moments_kernel(const uint8_t *__restrict__ img, int pitch, int width, int height,
                                                      int4 *__restrict__ mom1, mom2_t *__restrict__ mom2, int mom_pitch,
                                                      currx_t *__restrict__ currx) {
    const int x = blockIdx.x * MOM_THREADS + threadIdx.x;
    const int y0 = blockIdx.y * MOM_STRIP;
    const int y_end = min(y0 + MOM_STRIP, height - 8);  // positions y0 .. y_end-1 ; rows up to y+8 are read
    if (x > width - 16 || y0 >= y_end) return;  // x <= W-16: the 12-byte row reads stay inside the row
    const uint8_t *base = img + (size_t)y0 * pitch + x;
    const unsigned sh = (unsigned)(reinterpret_cast<uintptr_t>(base) & 3u) * 8u;
    moments_strip<false>(reinterpret_cast<const uint32_t *>(base - (sh >> 3)), pitch >> 2, sh, x, y0, y_end, width, height, mom1, mom2,
                         mom_pitch, currx);
}

// The same table from tiles staged in SHARED MEMORY by bulk asynchronous copies (cp.async.bulk + mbarrier, the TMA
// engine; sm_90+).  Why: this kernel is meant to run BESIDE the persistent ncc_kernel of the previous update, which
// leaves room for one small CTA per SM.  At that occupancy moments_kernel above is latency-bound (every row of the
// sliding window is a dependent global load): ~10x slower than alone, too slow to finish behind an ncc_kernel of an
// 8-GPU run (235 us per update at 4K), so the precompute landed on the critical path (26 ms of a 160 ms step).  Here
// one thread arms an mbarrier and issues one bulk copy per image row of the NEXT tile while the CTA computes the
// current tile from shared memory: the memory latency is paid once per tile and hidden by the double buffer, whatever
// the occupancy.  A tile = MB_COLS columns x MB_ROWS positions (+9 halo rows, +16 halo bytes); CTAs are persistent
// over the tiles.  Needs 16-byte aligned rows (pointer and pitch); dmf_api.cu falls back to moments_kernel otherwise.
constexpr int MB_COLS = 128, MB_ROWS = 32, MB_TILE_ROWS = MB_ROWS + 9, MB_ROWB = MB_COLS + 16, MB_STAGES = 2;
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_copy_g2s(void *dst, const void *src, unsigned bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src),
                 "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, unsigned parity) {
    unsigned done = 0;
    while (!done)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done)
                     : "r"(smem_u32(bar)), "r"(parity)
                     : "memory");
}

__global__ void __launch_bounds__(MB_COLS) moments_bulk_kernel(const uint8_t *__restrict__ img, int pitch, int width, int height,
                                                                int4 *__restrict__ mom1, mom2_t *__restrict__ mom2, int mom_pitch,
                                                                currx_t *__restrict__ currx, int tiles_x, int n_tiles) {
    __shared__ __align__(16) uint8_t tile[MB_STAGES][MB_TILE_ROWS * MB_ROWB];
    __shared__ __align__(8) uint64_t bar[MB_STAGES];
    const int tx = threadIdx.x;
    if (tx == 0) {
        for (int s = 0; s < MB_STAGES; ++s) mbar_init(&bar[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    // thread 0: arm the stage's barrier with the byte count of the tile and issue one bulk copy per image row
    auto issue = [&](int t, int s) {
        const int x0 = (t % tiles_x) * MB_COLS, y0 = (t / tiles_x) * MB_ROWS;
        const int rows = min(MB_TILE_ROWS, height - y0);
        const unsigned rb = (unsigned)min(MB_ROWB, pitch - x0);  // multiple of 16: pitch and x0 are
        mbar_arrive_expect_tx(&bar[s], rb * (unsigned)rows);
        const uint8_t *src = img + (size_t)y0 * pitch + x0;
        for (int r = 0; r < rows; ++r) bulk_copy_g2s(&tile[s][r * MB_ROWB], src + (size_t)r * pitch, rb, &bar[s]);
    };
    int it = 0;
    if (tx == 0 && (int)blockIdx.x < n_tiles) issue(blockIdx.x, 0);
    for (int t = blockIdx.x; t < n_tiles; t += gridDim.x, ++it) {
        const int s = it & 1;
        const int tn = t + gridDim.x;
        if (tx == 0 && tn < n_tiles) issue(tn, s ^ 1);  // stage s^1 was released by the barrier at the end of the previous pass
        mbar_wait(&bar[s], (unsigned)(it >> 1) & 1u);
        const int x0 = (t % tiles_x) * MB_COLS, y0 = (t / tiles_x) * MB_ROWS;
        const int x = x0 + tx;
        const int y_end = min(y0 + MB_ROWS, height - 8);
        if (x <= width - 16 && y0 < y_end) {
            const unsigned sh = (unsigned)(tx & 3) * 8u;
            moments_strip<true>(reinterpret_cast<const uint32_t *>(&tile[s][tx & ~3]), MB_ROWB >> 2, sh, x, y0, y_end, width, height, mom1,
                                mom2, mom_pitch, currx);
        }
        __syncthreads();  // every thread is done reading stage s before it is refilled two passes later
    }
}

// The integer part of one NCC (ref:449-480): everything that depends on the integer position (ix,iy) of
// the sample but not on its bilinear fractions — the four centred cross sums with the reference patch and
// the ten centred Gram sums of the block.  Consecutive samples of a search are 0.7 px apart, so about one
// in three (axis-aligned lines) falls on the same integer position as its predecessor and reuses these
// 14 integers without touching memory.
// R0lo/R0hi: reference rows as bytes (r0..r3),(r4,r5,r6,0) — they meet block columns 0..6 (window a=0);
// R1lo/R1hi: the same rows shifted by one byte, (0,r0,r1,r2),(r3..r6) — they meet block columns 1..7
// (window a=1) of the SAME unshifted block words, so the block needs no per-sample byte shifts.
struct SampleInts {
    int cR00, cR10, cR01, cR11;                                                // 49*R - Sr*S per window
    int g0000, g1010, g0101, g1111, g0010, g0111, g0001, g1011, gD;  // 49*G - S*S' (gD: the two diagonal terms summed)
};
// The FP64 part: combination with the bilinear weights of ref:169-172 (fractions fx, fy, ref:167-168);
// den1 = 49*sum r^2 - (sum r)^2.  int -> double conversions run on the XU pipe (I2F.F64).
__device__ __forceinline__ double ncc_combine(const SampleInts &s, double den1, double fx, double fy) {
    const double gx = 1.0 - fx, gy = 1.0 - fy;
    // w00 = gx*gy, w10 = fx*gy, w01 = gx*fy, w11 = fx*fy are separable, so
    //   num  = gy (gx cR00 + fx cR10) + fy (gx cR01 + fx cR11)
    //   den2 = w^T G w = gy^2 (gx^2 q00 + fx^2 q10 + 2 gx fx H0) + fy^2 (gx^2 q01 + fx^2 q11 + 2 gx fx H1)
    //                    + 2 gy fy (gx^2 V0 + fx^2 V1 + gx fx (D1 + D2))
    // with q = squares, H / V / D = horizontal / vertical / diagonal neighbour products of the centred Gram matrix;
    // the two diagonal terms share one weight, so the table carries their (exact) integer sum.
    double n0 = gx * (double)s.cR00; n0 = fma(fx, (double)s.cR10, n0);
    double n1 = gx * (double)s.cR01; n1 = fma(fx, (double)s.cR11, n1);
    const double num = fma(fy, n1, gy * n0);
    const double A = gx * gx, B = gx * fx, C = fx * fx, D = gy * gy, E = gy * fy, F = fy * fy;
    const double B2 = B + B, E2 = E + E;
    double t0 = A * (double)s.g0000; t0 = fma(C, (double)s.g1010, t0); t0 = fma(B2, (double)s.g0010, t0);
    double t1 = A * (double)s.g0101; t1 = fma(C, (double)s.g1111, t1); t1 = fma(B2, (double)s.g0111, t1);
    double t2 = A * (double)s.g0001; t2 = fma(C, (double)s.g1011, t2); t2 = fma(B, (double)s.gD, t2);
    double den2 = D * t0; den2 = fma(F, t1, den2); den2 = fma(E2, t2, den2);
    const double dd = fma(den1, den2, NCC_EPS_INT);
    return num * rsqrt(dd);
}

// K2b: NCC over the work units.

// One sample's memory operands: the 8 rows of the 8x8 block (one aligned 64-bit word each, from the expanded frame)
// and 5 vectors of the moment table.  All loads are issued before the first use (memory-level parallelism).
struct RawSample {
    uint32_t lo[8], hi[8];
    int4 m00, m10, m01, m11;
    int md;   // cD1 + cD2
};
// WIDTH: image width as a compile-time constant (0 = use P.width).  With a constant width the eight row loads share
// one base address and differ by immediate offsets (no per-row address arithmetic, no row-pointer registers).
template <int WIDTH>
__device__ __forceinline__ void load_raw(const KParams &P, int ix, int iy, RawSample &r) {
    const unsigned W = WIDTH ? (unsigned)WIDTH : (unsigned)P.width;
    // one element offset for the three tables (their pitch is the image width; W*H < 2^31)
    const unsigned o = (unsigned)(iy - 3) * W + (unsigned)(ix - 3);
    const currx_t *xp = P.currx + o;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const uint2 q = __ldg(xp + (size_t)j * W);
        r.lo[j] = q.x; r.hi[j] = q.y;
    }
    const int4 *m1 = P.mom1 + o;
    r.m00 = __ldg(m1); r.m10 = __ldg(m1 + 1);
    r.m01 = __ldg(m1 + W); r.m11 = __ldg(m1 + W + 1);
    r.md = __ldg(P.mom2 + o);
}
// Measured and dropped (bit-identical results, slower): prefetching the next sample's lines (CCTL.PF1, +34 %:
// profiles/r02_ab_ncc_prefetch_and_moments.txt, commit 04641b5) and carrying moment-table entries between neighbouring
// integer positions in registers (the per-lane step-type branches serialise the loads, +28 %: profiles/r02_ab_ncc_carry.txt,
// commit 43954c3); two samples in flight per thread (operands of sample j+1 requested before sample j is reduced: 128-168
// registers, 12-16 warps/SM, +36..60 %: profiles/r02_ab_ncc_two_samples_in_flight.txt); an L2 persisting window on the
// moment tables (no change) and claiming the next grab early with an L2 prefetch of its records (+3 %:
// profiles/r02_ab_l2_persist_and_grab_ahead.txt; all three in commit c87ba67).  Throughput follows the number of resident
// warps: the kernel is bound by L1 wavefronts and by dependent-issue latency, not by DRAM latency or load instructions.
// cross sums with the reference patch (window (a,b) = block columns a..a+6, rows b..b+6) + exact int32 centring.
// SHIFT_BLOCK: the windows a = 1 (block columns 1..7) are formed by shifting the BLOCK row by one byte per sample
// (2 ALU ops per row) instead of holding a second, byte-shifted copy of the reference patch (14 registers less: 80
// instead of 96, i.e. 4 instead of 3 CTAs per SM).
// Selected per context: 4 CTAs per SM (80 registers, 24 warps) pay off where the tables exceed L2 (4K: -5 % on ncc_kernel),
// not at 1080p and below (profiles/r02_ab_ncc_shift_block.txt).
template <bool SHIFT_BLOCK>
__device__ __forceinline__ SampleInts reduce_raw(const RawSample &r, const uint32_t (&R0lo)[7], const uint32_t (&R0hi)[7],
                                                 const uint32_t (&R1lo)[7], const uint32_t (&R1hi)[7], int nSr) {
    int R00 = 0, R10 = 0, R01 = 0, R11 = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        if (SHIFT_BLOCK) {
            const uint32_t lo1 = __funnelshift_r(r.lo[j], r.hi[j], 8), hi1 = r.hi[j] >> 8;
            if (j > 0) {
                R01 = dp4(R0lo[j - 1], r.lo[j], dp4(R0hi[j - 1], r.hi[j], R01));
                R11 = dp4(R0lo[j - 1], lo1, dp4(R0hi[j - 1], hi1, R11));
            }
            if (j < 7) {
                R00 = dp4(R0lo[j], r.lo[j], dp4(R0hi[j], r.hi[j], R00));
                R10 = dp4(R0lo[j], lo1, dp4(R0hi[j], hi1, R10));
            }
            continue;
        }
        if (j > 0) {
            R01 = dp4(R0lo[j - 1], r.lo[j], dp4(R0hi[j - 1], r.hi[j], R01));
            R11 = dp4(R1lo[j - 1], r.lo[j], dp4(R1hi[j - 1], r.hi[j], R11));
        }
        if (j < 7) {
            R00 = dp4(R0lo[j], r.lo[j], dp4(R0hi[j], r.hi[j], R00));
            R10 = dp4(R1lo[j], r.lo[j], dp4(R1hi[j], r.hi[j], R10));
        }
    }
    SampleInts s;
    // exact centring in int32 (all terms < 2^31)
    s.cR00 = NCC_AREA * R00 + nSr * r.m00.x; s.cR10 = NCC_AREA * R10 + nSr * r.m10.x;
    s.cR01 = NCC_AREA * R01 + nSr * r.m01.x; s.cR11 = NCC_AREA * R11 + nSr * r.m11.x;
    s.g0000 = r.m00.y; s.g1010 = r.m10.y; s.g0101 = r.m01.y; s.g1111 = r.m11.y;
    s.g0010 = r.m00.z; s.g0111 = r.m01.z; s.g0001 = r.m00.w; s.g1011 = r.m10.w;
    s.gD = r.md;
    return s;
}
// integer part and fraction of a sample coordinate c (0 <= c < 2^31), ref:167-168.  Adding 2^52 with round-down
// leaves floor(c) in the low mantissa word: DADDs on the FP64 pipe instead of F2I + I2F on the eighth-rate XU pipe.
__device__ __forceinline__ void split_coord(double c, int &i, double &f) {
    const double t = __dadd_rd(c, 4503599627370496.0);
    i = __double2loint(t);
    f = c - (t - 4503599627370496.0);
}

// Descriptors of the GRAB/32 slices at cursor value g (warp-uniform): unit words and their common length.
__device__ __forceinline__ void fetch_units(const KParams &P, const unsigned (&counts)[CHUNK + 1], unsigned total, unsigned g,
                                            int lane, unsigned (&units)[GRAB / 32], int (&lens)[GRAB / 32]) {
#pragma unroll
    for (int sub = 0; sub < GRAB / 32; ++sub) {
        const unsigned w0 = g + sub * 32;  // warp-uniform slot base (32-aligned)
        // segment of this slice (uniform): unit length L, first slot, number of units
        int L = 0;
        unsigned start = 0, cnt = 0, acc = 0;
#pragma unroll
        for (int c = CHUNK; c >= 1; --c) {
            const unsigned padded = (counts[c] + 31u) & ~31u;
            if (w0 >= acc && w0 < acc + padded) { L = c; start = acc; cnt = counts[c]; }
            acc += padded;
        }
        const unsigned idx = w0 - start + lane;
        const bool valid = (w0 < total) && (idx < cnt);
        lens[sub] = valid ? L : 0;
        units[sub] = 0;
        if (valid) units[sub] = (L == CHUNK) ? __ldg(P.units_full + idx) : __ldg(P.units_tail + (size_t)(L - 1) * P.n_pix + idx);
    }
}

template <int WIDTH, bool SHIFT_BLOCK>
__global__ void __launch_bounds__(NCC_THREADS, SHIFT_BLOCK ? DMF_NCC_MIN_BLOCKS + 1 : DMF_NCC_MIN_BLOCKS) ncc_kernel(const __grid_constant__ KParams P) {
    const int lane = threadIdx.x & 31;
    // padded, concatenated lists: length CHUNK first, then CHUNK-1, ..., 1; each segment 32-aligned
    unsigned counts[CHUNK + 1];
    unsigned total = 0;
#pragma unroll
    for (int c = CHUNK; c >= 1; --c) {
        counts[c] = P.ctrl->count[c];
        total += (counts[c] + 31u) & ~31u;
    }
    unsigned my_evals = 0;

    for (;;) {
        unsigned g = 0;
        if (lane == 0) g = atomicAdd(&P.ctrl->cursor, (unsigned)GRAB);
        g = __shfl_sync(0xffffffffu, g, 0);
        if (g >= total) break;
        // descriptors of all GRAB/32 slices first: their latency overlaps instead of heading every slice
        unsigned units[GRAB / 32];
        int lens[GRAB / 32];
        fetch_units(P, counts, total, g, lane, units, lens);
#pragma unroll
        for (int sub = 0; sub < GRAB / 32; ++sub) {
            const int L = lens[sub];
            if (L == 0) continue;
            const unsigned unit = units[sub];
            const unsigned slot = unit >> CHUNK_BITS;
            const int k0 = (int)(unit & ((1u << CHUNK_BITS) - 1u)) * CHUNK;
            // one 64-byte record fetch
            const PixelRec *rec = P.rec + slot;
            const double2 pm = rec->pm, dir = rec->dir;
            const int4 hv = *reinterpret_cast<const int4 *>(&rec->half);
            const int4 xy = rec->xy;
            const double half = __hiloint2double(hv.y, hv.x);
            const int nSr = hv.z;
            const double den1 = (double)hv.w;
            const int x = xy.x, y = xy.y;

            // reference patch of (x,y) into registers: 7 aligned 64-bit words of the expanded reference frame
            uint32_t R0lo[7], R0hi[7], R1lo[7], R1hi[7];
            {
                const unsigned W = WIDTH ? (unsigned)WIDTH : (unsigned)P.width;
                const uint2 *rp = P.refx + (unsigned)(y - 3) * W + (unsigned)x;
#pragma unroll
                for (int j = 0; j < 7; ++j) {
                    const uint2 q = __ldg(rp + (size_t)j * W);
                    R0lo[j] = q.x; R0hi[j] = q.y;
                    if (SHIFT_BLOCK) { R1lo[j] = 0; R1hi[j] = 0; }  // unused: the block is shifted instead (reduce_raw)
                    else { R1lo[j] = q.x << 8; R1hi[j] = __funnelshift_l(q.x, q.y, 8); }
                }
            }

            double best_v = -1.0;  // ref:430
            int best_k = -1;
            int hix = -1, hiy = -1;  // integer position whose SampleInts are held
            SampleInts si{};
            // l of the unit's first sample: k0 additions of the step, as the reference accumulates them (ref:432);
            // inside the loop the position of sample j+1 is computed before the NCC of sample j (off the critical path)
            double sx, sy;
            double l = dmf_geom::sample_l_acc(half, P.step, k0);
            sx = __dadd_rn(pm.x, __dmul_rn(l, dir.x));  // ref:433, unfused like the reference build
            sy = __dadd_rn(pm.y, __dmul_rn(l, dir.y));
#pragma unroll 1
            for (int j = 0; j < L; ++j) {
                const double cx = sx, cy = sy;
                l = __dadd_rn(l, P.step);
                sx = __dadd_rn(pm.x, __dmul_rn(l, dir.x));
                sy = __dadd_rn(pm.y, __dmul_rn(l, dir.y));
                // inside() ref:222-224
                const bool ok = cx >= P.bd && cy >= P.bd && cx + P.bd < P.wd && cy + P.bd <= P.hd;
                if (!ok) continue;
                int ix, iy;
                double fx, fy;
                split_coord(cx, ix, fx);
                split_coord(cy, iy, fy);
                if (ix != hix || iy != hiy) {
                    RawSample raw;
                    load_raw<WIDTH>(P, ix, iy, raw);
                    si = reduce_raw<SHIFT_BLOCK>(raw, R0lo, R0hi, R1lo, R1hi, nSr);
                    hix = ix; hiy = iy;
                }
                const double v = ncc_combine(si, den1, fx, fy);
                ++my_evals;
                if (v > best_v) { best_v = v; best_k = k0 + j; }  // first strict maximum ref:438-441
            }
            if (best_k >= 0) atomicMax(&P.best[slot], ncc_key(best_v, best_k));
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) my_evals += __shfl_down_sync(0xffffffffu, my_evals, o);
    if (lane == 0 && my_evals) atomicAdd(&P.counters[1], (unsigned long long)my_evals);
}

// ----------------------------------------------------------------------------------------
// Accept test ref:443 + updateDepthFilter ref:482-567 of one slot of the frame being finished: key, record and the
// state setup read arrive in ONE round trip.  Writes the fused state to the maps and returns it in (mu, c2).
__device__ __forceinline__ bool fuse_slot(const KParams &P, unsigned slot, int &x, int &y, double &mu, double &c2,
                                          unsigned long long &key, D3 &f_ref_out, bool &have_ray) {
    key = P.best_fin[slot];
    const PixelRec *rec = P.rec_fin + slot;
    const double2 pm = rec->pm, dir = rec->dir;
    const double half = rec->half;
    const int4 xy = rec->xy;
    const double2 mc = P.state_fin[slot];
    mu = mc.x; c2 = mc.y;
    x = xy.x; y = xy.y;
    have_ray = false;
    const bool accepted = key_has_winner(key) && !(key_ncc(key) < P.ncc_thresh);  // ref:443; sentinel: nothing beat -1.0
    if (accepted) {
        const int k = key_index(key);
        const dmf_geom::Camera cam = camera_of(P);
        const dmf_geom::V2 pmv{pm.x, pm.y}, dv{dir.x, dir.y};
        const dmf_geom::V2 pt_curr = dmf_geom::sample_pos(pmv, dv, dmf_geom::sample_l_acc(half, P.step, k));  // best_px_curr ref:440
        const D3 f_ref = dmf_geom::unit_ray(cam, (double)x, (double)y);
        f_ref_out = f_ref; have_ray = true;
        // updateDepthFilter ref:482-567 in the reference's operation order (ColPivHouseholderQR restated)
        const dmf_geom::Fused fu = dmf_geom::fuse(cam, P.qi, P.ti, P.ti_norm, f_ref, pt_curr, dv, mu, c2, P.inverse_depth != 0);
        mu = fu.mu;       // ref:560-562
        c2 = fu.sigma2;   // ref:564
        P.depth[(size_t)y * P.state_pitch + x] = mu;
        P.cov2[(size_t)y * P.state_pitch + x] = c2;
    }
    return accepted;
}

// counters of the finished frame: warp ballots -> shared -> ONE global atomic per CTA and counter (same-address
// atomics from every warp serialise in L2 and were the critical path of this kernel)
__device__ __forceinline__ void count_frame(const KParams &P, unsigned n_fin, bool accepted) {
    __shared__ unsigned s_acc;
    if (threadIdx.x == 0) s_acc = 0;
    __syncthreads();
    const unsigned c = __popc(__ballot_sync(0xffffffffu, accepted));
    if ((threadIdx.x & 31) == 0 && c) atomicAdd(&s_acc, c);
    __syncthreads();
    if (threadIdx.x == 0) {
        atomicAdd(&P.counters[0], (unsigned long long)n_fin);
        if (s_acc) atomicAdd(&P.counters[2], (unsigned long long)s_acc);
    }
}

// K2c: finishes a frame on its own (same grid as setup_kernel, thread = slot of the CTA's range): accept test + fusion,
// in place on the maps.  Used when the maps are read before the next update (end of a sequence, strict drop-in mode,
// debug planes).
#ifndef DMF_FUSE_MIN_BLOCKS
#define DMF_FUSE_MIN_BLOCKS 1
#endif
__global__ void __launch_bounds__(TILE_PIX, DMF_FUSE_MIN_BLOCKS) fuse_kernel(const __grid_constant__ KParams P) {
    const unsigned cta = blockIdx.y * gridDim.x + blockIdx.x;
    if (cta == 0 && threadIdx.x < sizeof(Ctrl) / sizeof(unsigned)) reinterpret_cast<unsigned *>(P.ctrl_zero)[threadIdx.x] = 0;
    const unsigned n_fin = P.cta_fin[cta];
    if (n_fin == 0) return;  // CTA-uniform
    bool accepted = false;
    if (threadIdx.x < n_fin) {
        int x, y;
        double mu, c2;
        unsigned long long key;
        D3 ray;
        bool have_ray;
        accepted = fuse_slot(P, cta * TILE_PIX + threadIdx.x, x, y, mu, c2, key, ray, have_ray);
        if (P.write_flags) {
            const size_t o = (size_t)y * P.flags_pitch + x;
            P.flags[o] = (uint8_t)(1 | (accepted ? 2 : 0));
            P.dbg_ncc[o] = (float)key_ncc(key);
            const unsigned kb = key_has_winner(key) ? (unsigned)key_index(key) : 0xFFFFu;
            const int trips = P.dbg_n[o] & 0xFFFF;
            P.dbg_n[o] = (trips << 16) | (int)kb;
        }
    }
    count_frame(P, n_fin, accepted);
}

// K2d: finishes update k AND sets up update k+1 (same grid as setup_kernel, thread = slot of update k).  A pixel that
// fails the gate ref:366 is never written again, so it stays inactive for the rest of the sequence: only the slots of
// update k can be active in update k+1.  The fused state goes from registers straight into the next search geometry;
// the maps are written but not read, and the converged / diverged pixels cost nothing any more.  P.q/P.t: pose of
// update k+1 (setup part); P.qi/P.ti/P.ti_norm: inverse pose of update k (fusion part).
// 4 CTAs of 256 threads per SM (64 registers): the kernel is a long chain of dependent FP64 operations (IEEE divisions
// and square roots in the reference's order), i.e. latency-bound: measured 37.0 ms per 1080p step at 2 CTAs / SM (84
// registers), 28.3 at 3, 24.1 at 4, 23.5 at 5 (profiles/r02_ab_advance_kernel.txt)
#ifndef DMF_ADV_MIN_BLOCKS
#define DMF_ADV_MIN_BLOCKS 4
#endif
__global__ void __launch_bounds__(TILE_PIX, DMF_ADV_MIN_BLOCKS) advance_kernel(const __grid_constant__ KParams P) {
    const unsigned cta = blockIdx.y * gridDim.x + blockIdx.x;
    if (cta == 0 && threadIdx.x < sizeof(Ctrl) / sizeof(unsigned)) reinterpret_cast<unsigned *>(P.ctrl_zero)[threadIdx.x] = 0;
    const unsigned n_fin = P.cta_fin[cta];
    if (n_fin == 0) {  // CTA-uniform: nothing left in this tile
        if (threadIdx.x == 0) P.cta_cnt[cta] = 0;
        return;
    }
    const bool have = threadIdx.x < n_fin;
    bool accepted = false;
    PixelWork w;
    w.x = 0; w.y = 0; w.mu = 0; w.c2 = 0; w.have_ray = false;
    if (have) {
        unsigned long long key;
        accepted = fuse_slot(P, cta * TILE_PIX + threadIdx.x, w.x, w.y, w.mu, w.c2, key, w.f_ref, w.have_ray);
    }
    prepare_pixel(P, w, have);
    emit_pixel(P, w, accepted, n_fin);
}

// ----------------------------------------------------------------------------------------
// Self-test of the grouped divisions of dmf_geometry.h (quo2 / quo3: shared reciprocal, one range test per group)
// against __ddiv_rn on pseudo-random operands: exponents from 2^-140 to 2^140 (beyond the fast range on both sides),
// plus zeros, infinities, NaNs and denormals.  Counts quotients whose bits differ (NaN matches NaN).
__device__ __forceinline__ unsigned long long mix64(unsigned long long h) {
    h ^= h >> 33; h *= 0xff51afd7ed558ccdull; h ^= h >> 33; h *= 0xc4ceb9fe1a85ec53ull; h ^= h >> 33;
    return h;
}
__device__ __forceinline__ double test_operand(unsigned long long h) {
    const unsigned sel = (unsigned)(h >> 58);  // 6 bits
    if (sel == 0) return 0.0;
    if (sel == 1) return __longlong_as_double(0x7ff0000000000000ll);                      // inf
    if (sel == 2) return __longlong_as_double(0x7ff8000000000000ll);                      // NaN
    if (sel == 3) return __longlong_as_double((long long)(h & 0x000fffffffffffffull));    // denormal
    const long long e = 1023 - 140 + (long long)((h >> 40) % 281);
    const unsigned long long sign = (h >> 57) & 1ull;
    return __longlong_as_double((long long)((sign << 63) | ((unsigned long long)e << 52) | (h & 0x000fffffffffffffull)));
}
__device__ __forceinline__ bool same_bits(double a, double b) {
    return (a != a && b != b) || __double_as_longlong(a) == __double_as_longlong(b);
}
__global__ void __launch_bounds__(256) division_selftest_kernel(unsigned long long seed, unsigned long long n, unsigned long long *mismatches) {
    unsigned long long bad = 0;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (unsigned long long)gridDim.x * blockDim.x) {
        const unsigned long long h = mix64(seed + i * 0x9e3779b97f4a7c15ull);
        const double b = test_operand(mix64(h ^ 1)), a0 = test_operand(mix64(h ^ 2)), a1 = test_operand(mix64(h ^ 3)), a2 = test_operand(mix64(h ^ 4));
        double q0, q1, q2, p0, p1;
        dmf_geom::quo3(a0, a1, a2, b, q0, q1, q2);
        dmf_geom::quo2(a1, a2, b, p0, p1);
        bad += !same_bits(q0, __ddiv_rn(a0, b)) + !same_bits(q1, __ddiv_rn(a1, b)) + !same_bits(q2, __ddiv_rn(a2, b)) +
               !same_bits(p0, __ddiv_rn(a1, b)) + !same_bits(p1, __ddiv_rn(a2, b));
    }
    if (bad) atomicAdd(mismatches, bad);
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

// ----------------------------------------------------------------------------------------
// "next" row: getPointCloudFromImageAndDistance, utils/pointcloud/pointcloud_from_image_depth.h:42-89,
// as called at ref:296-300: mask = variance > max_variance ? 0 : 255 (getMaskFromVariance ref:199-204), distance =
// the depth map (|OP| along the ray), T = identity.  Points come out in the reference's scan order (rows, then
// columns): pass 1 counts the valid pixels of every interior row, a one-block scan turns counts into offsets,
// pass 2 writes each row's points at its offset in column order.
__device__ __forceinline__ bool cloud_valid(double dist, double var, double max_variance) {
    return !(dist == 0) && !(var > max_variance);  // `distance == 0 || valid == 0 -> continue` (:66-67)
}

__global__ void __launch_bounds__(256) cloud_count_kernel(const double *__restrict__ dist, const double *__restrict__ var, int pitch,
                                                          int x0, int x1, const int *__restrict__ rows, double max_variance,
                                                          unsigned int *__restrict__ row_count) {
    const int y = rows[blockIdx.x];  // owned image rows in ascending order (a band, or the blocks of a cyclic context)
    unsigned n = 0;
    for (int x = x0 + threadIdx.x; x < x1; x += blockDim.x)
        n += cloud_valid(dist[(size_t)y * pitch + x], var[(size_t)y * pitch + x], max_variance) ? 1u : 0u;
    __shared__ unsigned s[8];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) n += __shfl_down_sync(0xffffffffu, n, o);
    if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = n;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned t = 0;
        for (int i = 0; i < 8; ++i) t += s[i];
        row_count[blockIdx.x] = t;
    }
}

// exclusive scan of n_rows counts in place (single block); total -> row_count[n_rows]
__global__ void __launch_bounds__(1024) cloud_scan_kernel(unsigned int *row_count, int n_rows) {
    __shared__ unsigned s[1024];
    __shared__ unsigned carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < n_rows; base += 1024) {
        const int i = base + threadIdx.x;
        const unsigned v = i < n_rows ? row_count[i] : 0u;
        s[threadIdx.x] = v;
        __syncthreads();
        for (int o = 1; o < 1024; o <<= 1) {
            const unsigned t = threadIdx.x >= o ? s[threadIdx.x - o] : 0u;
            __syncthreads();
            s[threadIdx.x] += t;
            __syncthreads();
        }
        if (i < n_rows) row_count[i] = carry + s[threadIdx.x] - v;
        __syncthreads();
        if (threadIdx.x == 1023) carry += s[1023];
        __syncthreads();
    }
    if (threadIdx.x == 0) row_count[n_rows] = carry;
}

__global__ void __launch_bounds__(256) cloud_write_kernel(const double *__restrict__ dist, const double *__restrict__ var, int pitch,
                                                          const uint8_t *__restrict__ color, int color_pitch, int channels,
                                                          int x0, int x1, const int *__restrict__ rows, double max_variance, double cx, double cy,
                                                          double fx, double fy, const unsigned int *__restrict__ row_offset,
                                                          float *__restrict__ xyz, uint8_t *__restrict__ rgb,
                                                          unsigned long long capacity) {
    const int y = rows[blockIdx.x];
    __shared__ unsigned s_warp[8];
    __shared__ unsigned s_run;
    if (threadIdx.x == 0) s_run = row_offset[blockIdx.x];
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int xb = x0; xb < x1; xb += blockDim.x) {
        const int x = xb + threadIdx.x;
        bool ok = false;
        double d = 0;
        if (x < x1) {
            d = dist[(size_t)y * pitch + x];
            ok = cloud_valid(d, var[(size_t)y * pitch + x], max_variance);
        }
        const unsigned bal = __ballot_sync(0xffffffffu, ok);
        if (lane == 0) s_warp[warp] = __popc(bal);
        __syncthreads();
        unsigned before = 0, total = 0;
        for (int w = 0; w < 8; ++w) { before += w < warp ? s_warp[w] : 0u; total += s_warp[w]; }
        if (ok) {
            const unsigned long long slot = (unsigned long long)s_run + before + __popc(bal & ((1u << lane) - 1u));
            if (slot < capacity) {
                // point = normalize((u-cx)/fx, (v-cy)/fy, 1) * distance   (:68-73), stored as float like PointXYZRGB
                // (no FMA contraction: the floats are bit-identical to the reference's g++ build)
                double px = ((double)x - cx) / fx, py = ((double)y - cy) / fy, pz = 1.0;
                const double z = __dadd_rn(__dmul_rn(px, px), __dadd_rn(__dmul_rn(py, py), __dmul_rn(pz, pz)));
                if (z > 0) { const double nrm = sqrt(z); px /= nrm; py /= nrm; pz /= nrm; }
                xyz[3 * slot + 0] = (float)__dmul_rn(px, d); xyz[3 * slot + 1] = (float)__dmul_rn(py, d); xyz[3 * slot + 2] = (float)__dmul_rn(pz, d);
                const uint8_t *c = color + (size_t)y * color_pitch + (size_t)x * channels;
                rgb[3 * slot + 0] = channels >= 3 ? c[2] : c[0];  // r (:79-81: b,g,r = data[0..2])
                rgb[3 * slot + 1] = channels >= 3 ? c[1] : c[0];
                rgb[3 * slot + 2] = c[0];
            }
        }
        __syncthreads();
        if (threadIdx.x == 0) s_run += total;
        __syncthreads();
    }
}

}  // namespace dmf
