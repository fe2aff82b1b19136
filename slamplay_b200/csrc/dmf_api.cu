// dmf_api.cu — C ABI (include/dmf.h) over the sm_100a kernels in dmf_kernels.cuh.
// Host-side logic only: context, HBM-resident state, double-buffered frame upload,
// pose inversion (Sophus SE3::inverse, used at ref:491), launches.
#include "../../include/dmf.h"
#include "dmf_internal.h"
#include "dmf_kernels.cuh"

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <vector>

namespace {

thread_local std::string g_err;

struct dmf_ctx_impl {
    dmf_params prm{};
    int device = 0;
    // rows owned: local row rl -> image row  row0 + ((rl / blk) * cyc + ph) * blk + rl % blk
    int row0 = 0, blk = 1, cyc = 1, ph = 0, n_rows = 0;
    int rev_round = -1;  // incomplete last round of a block-cyclic dealing (dealt in reverse), or -1
    std::vector<std::pair<int, int>> spans;     // owned interior rows as ascending [y0, y1) intervals
    std::vector<std::pair<int, int>> io_spans;  // rows moved by upload / download (spans, plus border rows a contiguous band asked for)
    // stream: advance (fusion + set-up) -> ncc of every update; mom_stream: moments_kernel of the NEXT update runs beside them
    cudaStream_t stream = nullptr, copy_stream = nullptr, mom_stream = nullptr;
    // images
    uint8_t *d_ref = nullptr;
    uint8_t *d_curr[2] = {nullptr, nullptr};
    uint8_t *h_stage[2] = {nullptr, nullptr};
    cudaEvent_t ev_copied[2] = {nullptr, nullptr};   // H2D of buffer b finished (copy stream)
    cudaEvent_t ev_consumed[2] = {nullptr, nullptr}; // kernel that read buffer b finished (compute stream)
    cudaEvent_t ev_ext = nullptr;
    int img_pitch = 0;  // bytes, multiple of 16
    int2 *d_refstat = nullptr;
    double *d_depth = nullptr, *d_cov2 = nullptr, *d_truth = nullptr;
    uint8_t *d_flags = nullptr, *d_mask = nullptr;
    float *d_dbg_ncc = nullptr;
    int *d_dbg_n = nullptr;
    unsigned long long *d_counters = nullptr;  // 3 counters + eval (sum_sq as double bits, count)
    double *d_eval = nullptr;                  // [0] = sum_sq ; count lives in d_counters[3]
    // per-update scratch of the advance / setup -> ncc -> fuse pipeline
    // slot-indexed scratch, double-buffered by update parity (advance_kernel reads update k's while writing k+1's)
    dmf::PixelRec *d_rec[2] = {nullptr, nullptr};            // n_pix 64-byte records
    unsigned long long *d_best[2] = {nullptr, nullptr};
    unsigned int *d_units_full = nullptr, *d_units_tail = nullptr;
    dmf::Ctrl *d_ctrl = nullptr;               // three control blocks: update u uses [u % 3]; the kernel that finishes update f re-arms [(f + 2) % 3]
    double2 *d_state_c[2] = {nullptr, nullptr};  // per slot: (depth, cov2) the update started from
    unsigned long long seq = 0;                // index of the next update
    unsigned int *d_cta_cnt = nullptr;         // two arrays (update parity) of per-CTA active-pixel counts
    int tiles_x = 0, n_bands = 0, n_ctas = 0, n_slots = 0;
    // Lazy fusion: the fusion of the last update runs inside the NEXT update's advance_kernel, or in fuse_kernel as
    // soon as anything reads or replaces the maps (flush_pending).
    bool pending = false;
    double pend_qi[4] = {0, 0, 0, 1}, pend_ti[3] = {0, 0, 0}, pend_ti_norm = 0;
    // per-frame block-moment table + expanded current frame (moments_kernel), double-buffered by frame parity
    int4 *d_mom1[2] = {nullptr, nullptr};
    dmf::mom2_t *d_mom2[2] = {nullptr, nullptr};
    dmf::currx_t *d_currx[2] = {nullptr, nullptr};
    cudaEvent_t ev_mom_done[2] = {nullptr, nullptr};  // moments_kernel wrote table b (mom_stream)
    cudaEvent_t ev_tab_free[2] = {nullptr, nullptr};  // ncc_kernel that read table b finished (stream)
    cudaEvent_t ev_adv_done[2] = {nullptr, nullptr};  // advance / setup kernel of update u finished (stream), by parity of u
    bool mom_gate = true;                             // moments(u) is released with ncc(u-1), not earlier (DMF_MOMENTS_GATE=0: off)
    cudaEvent_t ev_frame = nullptr;                   // the frame of this update is complete in HBM
    uint2 *d_refx = nullptr;                   // expanded reference frame (ref_expand_kernel)
    int n_pix = 0, ncc_grid = 0;
    bool mom_bulk = true;                      // moments_bulk_kernel (DMF_MOMENTS=legacy selects moments_kernel: A/B runs)
    // DMF_TRACE=<file> (diagnostic): %globaltimer stamps on the streams around the three kernels of every update,
    // written to <file> by dmf_destroy: when does moments(u) run relative to ncc(u-1)?
    unsigned long long *d_trace = nullptr;
    int trace_cap = 0;
    std::string trace_path;
    int mom_repeat = 1;                        // DMF_MOMENTS_REPEAT (diagnostic): launches per update; the extra time per update is the exposed cost of one
    int mom_grid = 296;                        // persistent CTAs of moments_bulk_kernel (DMF_MOMENTS_CTAS_PER_SM x SMs)
    void (*ncc_fn)(dmf::KParams) = nullptr;    // ncc_kernel specialised for the image width (BASELINE.json's resolutions) or generic
    // optional per-kernel timing (dmf_set_timing): 5 events per update bracket the 4 timing slots
    bool timing_on = false;
    std::vector<cudaEvent_t> ev_pool;
    size_t ev_used = 0;
    std::vector<cudaEvent_t> ev_flush;         // pairs around the stand-alone fuse_kernel launched by flush_pending (timing on)
    size_t ev_flush_used = 0;
    double timing_ms[4] = {0, 0, 0, 0};
    unsigned long long timing_frames = 0;
    bool have_ref = false, flags_on = false, have_truth = false;
    // strict drop-in mode (dmf_update_strict): content hash of the reference image on the device, pinned shadow copies of
    // the maps as last downloaded (an unchanged host map is not uploaded again)
    unsigned long long ref_hash = 0;
    bool ref_hash_valid = false, shadow_valid = false;
    double *h_shadow[2] = {nullptr, nullptr};
    unsigned long long frames = 0;
    unsigned long long frame_idx = 0;
    std::string err;
};

int fail(dmf_ctx_impl *c, int code, const std::string &msg) {
    g_err = msg;
    if (c) c->err = msg;
    return code;
}

#define CU(call)                                                                                       \
    do {                                                                                               \
        cudaError_t e_ = (call);                                                                       \
        if (e_ != cudaSuccess)                                                                         \
            return fail(c, DMF_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_));          \
    } while (0)

struct Q4 { double x, y, z, w; };
void rot(const Q4 &q, const double v[3], double out[3]) {  // Eigen _transformVector
    double ux = q.y * v[2] - q.z * v[1], uy = q.z * v[0] - q.x * v[2], uz = q.x * v[1] - q.y * v[0];
    ux += ux; uy += uy; uz += uz;
    out[0] = v[0] + q.w * ux + (q.y * uz - q.z * uy);
    out[1] = v[1] + q.w * uy + (q.z * ux - q.x * uz);
    out[2] = v[2] + q.w * uz + (q.x * uy - q.y * ux);
}

// Sophus SE3::inverse(): invR = SO3(conj(q)) (constructor normalises), t' = invR * (t * -1)
void se3_inverse(const double q[4], const double t[3], double qi[4], double ti[3]) {
    Q4 c{-q[0], -q[1], -q[2], q[3]};
    double n = std::sqrt((c.x * c.x + c.z * c.z) + (c.y * c.y + c.w * c.w));
    c.x /= n; c.y /= n; c.z /= n; c.w /= n;
    double nt[3] = {t[0] * -1.0, t[1] * -1.0, t[2] * -1.0};
    rot(c, nt, ti);
    qi[0] = c.x; qi[1] = c.y; qi[2] = c.z; qi[3] = c.w;
}

int check_params(const dmf_params *p, std::string &why) {
    if (!p) { why = "params is NULL"; return -1; }
    if (p->width < 64 || p->height < 64 || p->width > 32768 || p->height > 32768) { why = "width/height out of range [64,32768]"; return -1; }
    if (p->ncc_half != 3) { why = "only ncc_half == 3 (7x7 window, ref:79) is supported"; return -1; }
    if (p->border < 13 || 2 * p->border >= p->width || 2 * p->border >= p->height) { why = "border must be >= 13 (the moment table covers block positions x <= width-16, y <= height-9) and < min(width,height)/2"; return -1; }
    // the arg-max key holds sample indices < 510 (KEY_IDX_BITS = 9); the work-unit encoding chunk indices < 64
    if (!(p->step > 0) || !(p->max_half_len >= 0) || !(p->max_half_len / p->step <= 250.0)) { why = "step must be > 0 and max_half_len/step <= 250"; return -1; }
    if (!(p->fx != 0) || !(p->fy != 0)) { why = "fx, fy must be non-zero"; return -1; }
    if (!(p->min_cov < p->max_cov)) { why = "min_cov must be < max_cov"; return -1; }
    return 0;
}

void fill_kparams(dmf_ctx_impl *c, dmf::KParams &K, unsigned long long u) {
    const dmf_params &p = c->prm;
    const int b = (int)(u & 1);  // parity of update u: slot buffers and moment-table buffer
    K.width = p.width; K.height = p.height; K.border = p.border;
    K.row0 = c->row0; K.blk = c->blk; K.cyc = c->cyc; K.ph = c->ph; K.n_rows = c->n_rows; K.rev_round = c->rev_round;
    K.inverse_depth = p.inverse_depth; K.write_flags = c->flags_on ? 1 : 0;
    K.ncc_thresh = p.ncc_thresh;
    K.fx = p.fx; K.fy = p.fy; K.cx = p.cx; K.cy = p.cy;
    K.step = p.step; K.max_half_len = p.max_half_len; K.min_depth = p.min_depth; K.n_sigma = p.n_sigma;
    K.min_cov = p.min_cov; K.max_cov = p.max_cov;
    K.bd = (double)p.border; K.wd = (double)p.width; K.hd = (double)p.height;
    K.ref = c->d_ref; K.refx = c->d_refx; K.refstat = c->d_refstat;
    K.depth = c->d_depth; K.cov2 = c->d_cov2; K.flags = c->d_flags; K.dbg_ncc = c->d_dbg_ncc; K.dbg_n = c->d_dbg_n; K.counters = c->d_counters;
    K.ref_pitch = c->img_pitch; K.stat_pitch = p.width; K.state_pitch = p.width;
    K.flags_pitch = p.width;
    K.wi = p.width - 2 * p.border;
    K.n_pix = c->n_pix;
    K.mom1 = c->d_mom1[b]; K.mom2 = c->d_mom2[b]; K.mom_pitch = p.width; K.currx = c->d_currx[b];
    K.units_full = c->d_units_full; K.units_tail = c->d_units_tail;
    K.rec = c->d_rec[b]; K.state_c = c->d_state_c[b]; K.best = c->d_best[b];
    K.ctrl = c->d_ctrl + (u % 3);
    K.cta_cnt = c->d_cta_cnt + (size_t)b * c->n_ctas;
}

// the update that is finished by a kernel launched with K: update f (its slot buffers, control block, inverse pose)
void fill_finish(dmf_ctx_impl *c, dmf::KParams &K, unsigned long long f) {
    const int b = (int)(f & 1);
    K.rec_fin = c->d_rec[b]; K.state_fin = c->d_state_c[b]; K.best_fin = c->d_best[b];
    K.cta_fin = c->d_cta_cnt + (size_t)b * c->n_ctas;
    K.ctrl_zero = c->d_ctrl + ((f + 2) % 3);
    for (int i = 0; i < 4; ++i) K.qi[i] = c->pend_qi[i];
    for (int i = 0; i < 3; ++i) K.ti[i] = c->pend_ti[i];
    K.ti_norm = c->pend_ti_norm;
}

// Runs the fusion of the last update if it is still pending (the maps are about to be read or replaced).
int flush_pending(dmf_ctx_impl *c) {
    if (!c->pending) return DMF_OK;
    CU(cudaSetDevice(c->device));
    dmf::KParams K{};
    fill_kparams(c, K, c->seq - 1);
    fill_finish(c, K, c->seq - 1);
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    if (c->timing_on) {  // the deferred fusion of the last update of a sequence belongs to the per-kernel split as well
        while (c->ev_flush.size() < c->ev_flush_used + 2) {
            cudaEvent_t e;
            CU(cudaEventCreate(&e));
            c->ev_flush.push_back(e);
        }
        e0 = c->ev_flush[c->ev_flush_used++];
        e1 = c->ev_flush[c->ev_flush_used++];
        CU(cudaEventRecord(e0, c->stream));
    }
    dmf::fuse_kernel<<<dim3(c->tiles_x, c->n_bands), dmf::TILE_PIX, 0, c->stream>>>(K);
    if (e1) CU(cudaEventRecord(e1, c->stream));
    CU(cudaGetLastError());
    c->pending = false;
    return DMF_OK;
}

__global__ void stamp_kernel(unsigned long long *dst) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    *dst = t;
}

// frame_ready: event after which the frame at d_curr is complete (NULL: it already is when this call is made).
int launch_update(dmf_ctx_impl *c, const uint8_t *d_curr, int curr_pitch, const double q[4], const double t[3],
                  cudaEvent_t frame_ready, cudaEvent_t frame_consumed) {
    const dmf_params &p = c->prm;
    const int rows = c->n_rows;
    if (rows > 0) {
        // debug planes describe ONE update: with them on, every update is set up from the maps and fused at once
        if (c->pending && c->flags_on) { int rc = flush_pending(c); if (rc) return rc; }
        const unsigned long long u = c->seq;
        const int b = (int)(u & 1);
        dmf::KParams K{};
        fill_kparams(c, K, u);
        K.curr = d_curr; K.curr_pitch = curr_pitch;
        for (int i = 0; i < 4; ++i) K.q[i] = q[i];
        for (int i = 0; i < 3; ++i) K.t[i] = t[i];
        const bool merged = c->pending;
        if (merged) fill_finish(c, K, u - 1);  // inverse pose of update u-1 for the fusion part
        dim3 grid((K.wi + dmf::TILE_W - 1) / dmf::TILE_W, (rows + dmf::TILE_H - 1) / dmf::TILE_H);
        dim3 mgrid((p.width - 15 + dmf::MOM_THREADS - 1) / dmf::MOM_THREADS, (p.height - 8 + dmf::MOM_STRIP - 1) / dmf::MOM_STRIP);
        cudaEvent_t ev[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
        if (c->timing_on) {
            for (int i = 0; i < 5; ++i) {
                if (c->ev_used == c->ev_pool.size()) {
                    cudaEvent_t e;
                    CU(cudaEventCreate(&e));
                    c->ev_pool.push_back(e);
                }
                ev[i] = c->ev_pool[c->ev_used++];
            }
        }
        // moments_kernel needs only the frame, not the state: it runs on its own stream, beside setup / ncc / fuse of
        // the PREVIOUS update (whose ncc_kernel reads the other table buffer).  With per-kernel timing on, everything
        // is serialised on the context stream so that the event pairs bracket one kernel each.
        cudaStream_t ms = c->timing_on ? c->stream : c->mom_stream;
        if (frame_ready) CU(cudaStreamWaitEvent(ms, frame_ready, 0));
        if (ms != c->stream) {
            CU(cudaStreamWaitEvent(ms, c->ev_tab_free[b], 0));  // ncc_kernel of two updates ago has released the table buffer
            // ... and not before the ncc_kernel of the PREVIOUS update is released: beside advance_kernel (which fills every
            // register file) the precompute would only take its place in the queue; beside ncc_kernel it runs in the
            // registers that kernel leaves free
            if (c->mom_gate && u > 0) CU(cudaStreamWaitEvent(ms, c->ev_adv_done[b ^ 1], 0));
        }
        unsigned long long *tr = (c->d_trace && u < (unsigned long long)c->trace_cap) ? c->d_trace + 6 * u : nullptr;
        if (tr) { stamp_kernel<<<1, 1, 0, c->stream>>>(tr + 0); stamp_kernel<<<1, 1, 0, ms>>>(tr + 2); }
        if (ev[0]) CU(cudaEventRecord(ev[0], c->stream));
        if (merged) dmf::advance_kernel<<<grid, dmf::TILE_PIX, 0, c->stream>>>(K);  // fusion of u-1 + setup of u
        else dmf::setup_kernel<<<grid, dmf::TILE_PIX, 0, c->stream>>>(K);
        if (ev[1]) CU(cudaEventRecord(ev[1], c->stream));
        if (tr) stamp_kernel<<<1, 1, 0, c->stream>>>(tr + 1);
        CU(cudaEventRecord(c->ev_adv_done[b], c->stream));
        for (int rep = 0; rep < c->mom_repeat; ++rep)
        if (c->mom_bulk && (reinterpret_cast<uintptr_t>(d_curr) & 15u) == 0 && (curr_pitch & 15) == 0) {
            // tiles staged in shared memory by bulk asynchronous copies: efficient at the one-CTA-per-SM occupancy that is
            // left beside the persistent ncc_kernel of the previous update
            const int tiles_x = (p.width - 15 + dmf::MB_COLS - 1) / dmf::MB_COLS, tiles_y = (p.height - 8 + dmf::MB_ROWS - 1) / dmf::MB_ROWS;
            const int n_tiles = tiles_x * tiles_y;
            const int grid_b = n_tiles < c->mom_grid ? n_tiles : c->mom_grid;
            dmf::moments_bulk_kernel<<<grid_b, dmf::MB_COLS, 0, ms>>>(d_curr, curr_pitch, p.width, p.height, c->d_mom1[b], c->d_mom2[b], p.width,
                                                                      c->d_currx[b], tiles_x, n_tiles);
        } else {
            dmf::moments_kernel<<<mgrid, dmf::MOM_THREADS, 0, ms>>>(d_curr, curr_pitch, p.width, p.height, c->d_mom1[b], c->d_mom2[b], p.width, c->d_currx[b]);
        }
        if (tr) stamp_kernel<<<1, 1, 0, ms>>>(tr + 3);
        if (frame_consumed) CU(cudaEventRecord(frame_consumed, ms));
        if (ms != c->stream) {
            CU(cudaEventRecord(c->ev_mom_done[b], ms));
            CU(cudaStreamWaitEvent(c->stream, c->ev_mom_done[b], 0));
        }
        if (ev[2]) CU(cudaEventRecord(ev[2], c->stream));
        if (tr) stamp_kernel<<<1, 1, 0, c->stream>>>(tr + 4);
        c->ncc_fn<<<c->ncc_grid, dmf::NCC_THREADS, 0, c->stream>>>(K);
        if (tr) stamp_kernel<<<1, 1, 0, c->stream>>>(tr + 5);
        CU(cudaEventRecord(c->ev_tab_free[b], c->stream));
        if (ev[3]) CU(cudaEventRecord(ev[3], c->stream));
        CU(cudaGetLastError());
        // the fusion of this update stays pending: the next update's advance_kernel or flush_pending() runs it
        se3_inverse(q, t, c->pend_qi, c->pend_ti);
        c->pend_ti_norm = std::sqrt(c->pend_ti[0] * c->pend_ti[0] + (c->pend_ti[1] * c->pend_ti[1] + c->pend_ti[2] * c->pend_ti[2]));
        c->pending = true;
        c->seq++;
        if (c->flags_on) { int rc = flush_pending(c); if (rc) return rc; }
        if (ev[4]) CU(cudaEventRecord(ev[4], c->stream));
    } else if (frame_consumed) {
        CU(cudaEventRecord(frame_consumed, c->stream));
    }
    c->frames++;
    return DMF_OK;
}

}  // namespace

struct dmf_ctx : dmf_ctx_impl {};

void dmf_internal_set_error(const char *msg) { g_err = msg ? msg : ""; }

int dmf_internal_stage_begin(dmf_ctx *c, int width, int height, dmf_internal_stage *st) {
    if (!c || !st) return fail(c, DMF_ERR_INVALID, "stage: NULL argument");
    if (!c->have_ref) return fail(c, DMF_ERR_STATE, "dmf_update_ring: dmf_set_reference() has not been called");
    if (width != c->prm.width || height != c->prm.height) return fail(c, DMF_ERR_INVALID, "dmf_update_ring: the ring's frame size differs from the context's");
    CU(cudaSetDevice(c->device));
    const int b = (int)(c->frame_idx & 1);
    c->frame_idx++;
    CU(cudaStreamWaitEvent(c->copy_stream, c->ev_consumed[b], 0));
    st->copy_stream = (void *)c->copy_stream;
    st->dst = c->d_curr[b];
    st->pitch = c->img_pitch;
    st->buffer = b;
    return DMF_OK;
}

int dmf_internal_stage_launch(dmf_ctx *c, const dmf_internal_stage *st, const double q[4], const double t[3]) {
    const int b = st->buffer;
    CU(cudaEventRecord(c->ev_copied[b], c->copy_stream));
    return launch_update(c, c->d_curr[b], c->img_pitch, q, t, c->ev_copied[b], c->ev_consumed[b]);
}

extern "C" {

int dmf_abi_version(void) { return DMF_ABI_VERSION; }

const char *dmf_build_info(void) { return "slamplay_b200 dmf: sm_100a, built " __DATE__ " " __TIME__; }

const char *dmf_last_error(const dmf_ctx *ctx) { return ctx ? ctx->err.c_str() : g_err.c_str(); }

int dmf_default_params(dmf_params *p, int width, int height, int inverse_depth) {
    dmf_ctx_impl *c = nullptr;
    if (!p || width <= 0 || height <= 0) return fail(c, DMF_ERR_INVALID, "dmf_default_params: bad arguments");
    std::memset(p, 0, sizeof(*p));
    p->width = width; p->height = height;
    p->border = 20;    // ref:72
    p->ncc_half = 3;   // ref:79
    if (width == 640 && height == 480) {  // ref:73-78, float literals widened to double
        p->fx = 481.2f; p->fy = -480.0f; p->cx = 319.5f; p->cy = 239.5f;
    } else {
        const double s = (double)width / 640.0;
        p->fx = (double)481.2f * s; p->fy = -480.0 * s;
        p->cx = 0.5 * (width - 1); p->cy = 0.5 * (height - 1);
    }
    p->step = 0.7; p->max_half_len = 100; p->min_depth = 0.1; p->n_sigma = 3;  // ref:432,422,414,412
    p->ncc_thresh = 0.85f;                                                     // ref:443
    if (inverse_depth) { p->min_cov = 0.0001; p->max_cov = 1; }                // ref:82-83
    else { const double good_error = 0.01; p->min_cov = good_error * good_error; p->max_cov = 10; }  // ref:85-87
    p->inverse_depth = inverse_depth ? 1 : 0;
    return DMF_OK;
}

static int create_common(const dmf_params *params, int device, int row_begin, int row_end, int block_rows, int n_parts,
                         int part, dmf_ctx **out) {
    dmf_ctx_impl *c = nullptr;
    if (!out) return fail(c, DMF_ERR_INVALID, "dmf_create: out is NULL");
    *out = nullptr;
    std::string why;
    if (check_params(params, why)) return fail(c, DMF_ERR_INVALID, "dmf_create: " + why);
    if (row_begin < 0 || row_end > params->height || row_begin > row_end)
        return fail(c, DMF_ERR_INVALID, "dmf_create: need 0 <= row_begin <= row_end <= height");
    if (block_rows < 1 || n_parts < 1 || part < 0 || part >= n_parts)
        return fail(c, DMF_ERR_INVALID, "dmf_create_cyclic: need block_rows >= 1 and 0 <= part < n_parts");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(c, DMF_ERR_CUDA, std::string("dmf_create: no CUDA device (") + cudaGetErrorString(e) + "); there is no CPU fallback");
    if (device < 0 || device >= ndev) return fail(c, DMF_ERR_INVALID, "dmf_create: device index out of range");
    cudaDeviceProp prop{};
    CU(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10)
        return fail(c, DMF_ERR_CUDA, std::string("dmf_create: device '") + prop.name + "' is not sm_100 (the kernels are built for sm_100a only)");
    CU(cudaSetDevice(device));

    dmf_ctx *ctx = new (std::nothrow) dmf_ctx();
    if (!ctx) return fail(c, DMF_ERR_NOMEM, "dmf_create: out of host memory");
    c = ctx;
    c->prm = *params;
    c->device = device;
    {
        const int lo = params->border, hi = params->height - params->border;
        if (n_parts == 1) {  // contiguous band
            int y0 = row_begin < lo ? lo : row_begin, y1 = row_end > hi ? hi : row_end;
            if (y1 < y0) y1 = y0;
            c->row0 = y0; c->blk = (y1 - y0) > 0 ? (y1 - y0) : 1; c->cyc = 1; c->ph = 0; c->n_rows = y1 - y0;
            if (y1 > y0) c->spans.push_back({y0, y1});
            if (row_end > row_begin) c->io_spans.push_back({row_begin, row_end});
        } else {             // block-cyclic: blocks of block_rows interior rows dealt round-robin
            c->row0 = lo; c->blk = block_rows; c->cyc = n_parts; c->ph = part; c->n_rows = 0;
            // round k deals blocks k*n_parts .. k*n_parts + n_parts-1; odd rounds in reverse order (boustrophedon).
            // Only the last block of the image can be short, and a context's rounds stop at its first missing block.
            // An incomplete last round is dealt from the highest context down whatever its parity: the rows next to
            // the image border converge last (their matches leave the frame), so the context that holds the first
            // block should not also receive an extra last one (measured at 1080p / 8 contexts: 1.20x -> the mean load).
            const int n_blocks = (hi - lo + block_rows - 1) / block_rows;
            c->rev_round = (n_blocks % n_parts) ? n_blocks / n_parts : -1;
            for (int k = 0;; ++k) {
                const int pos = ((k & 1) || k == c->rev_round) ? (n_parts - 1 - part) : part;
                const int y0 = lo + (k * n_parts + pos) * block_rows;
                if (y0 >= hi) break;
                const int y1 = (y0 + block_rows < hi) ? y0 + block_rows : hi;
                c->spans.push_back({y0, y1});
                c->n_rows += y1 - y0;
            }
            c->io_spans = c->spans;
        }
    }
    const size_t W = params->width, H = params->height;
    c->img_pitch = (int)((W + 15) / 16 * 16);
#define CUX(call)                                                                                         \
    do {                                                                                                  \
        cudaError_t e2_ = (call);                                                                         \
        if (e2_ != cudaSuccess) {                                                                         \
            int rc_ = fail(nullptr, DMF_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e2_));   \
            dmf_destroy(ctx);                                                                             \
            return rc_;                                                                                   \
        }                                                                                                 \
    } while (0)
    {
        // The frame-only precompute of the NEXT update is released together with the persistent ncc_kernel of the current
        // one (see launch_update) and must only fill the room ncc_kernel leaves (one small CTA per SM): the context
        // stream gets the higher priority, so that all ncc_kernel CTAs are placed first.
        int prio_lo = 0, prio_hi = 0;
        CUX(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
        CUX(cudaStreamCreateWithPriority(&c->stream, cudaStreamNonBlocking, prio_hi));
        CUX(cudaStreamCreateWithPriority(&c->copy_stream, cudaStreamNonBlocking, prio_hi));
        CUX(cudaStreamCreateWithPriority(&c->mom_stream, cudaStreamNonBlocking, prio_lo));
    }
    CUX(cudaEventCreateWithFlags(&c->ev_frame, cudaEventDisableTiming));
    const size_t img_bytes = (size_t)c->img_pitch * H;
    CUX(cudaMalloc(&c->d_ref, img_bytes));
    for (int b = 0; b < 2; ++b) {
        CUX(cudaMalloc(&c->d_curr[b], img_bytes));
        CUX(cudaMallocHost(&c->h_stage[b], img_bytes));
        CUX(cudaEventCreateWithFlags(&c->ev_copied[b], cudaEventDisableTiming));
        CUX(cudaEventCreateWithFlags(&c->ev_consumed[b], cudaEventDisableTiming));
    }
    CUX(cudaEventCreateWithFlags(&c->ev_ext, cudaEventDisableTiming));
    CUX(cudaMalloc(&c->d_refstat, W * H * sizeof(int2)));
    CUX(cudaMalloc(&c->d_depth, W * H * sizeof(double)));
    CUX(cudaMalloc(&c->d_cov2, W * H * sizeof(double)));
    CUX(cudaMalloc(&c->d_flags, W * H));
    CUX(cudaMalloc(&c->d_counters, 4 * sizeof(unsigned long long)));
    CUX(cudaMalloc(&c->d_eval, sizeof(double)));
    CUX(cudaMemsetAsync(c->d_counters, 0, 4 * sizeof(unsigned long long), c->stream));
    CUX(cudaMemsetAsync(c->d_flags, 0, W * H, c->stream));
    CUX(cudaMemsetAsync(c->d_refstat, 0, W * H * sizeof(int2), c->stream));
    CUX(cudaMemsetAsync(c->d_depth, 0, W * H * sizeof(double), c->stream));
    CUX(cudaMemsetAsync(c->d_cov2, 0, W * H * sizeof(double), c->stream));
    {
        // scratch of the advance / setup -> ncc -> fuse pipeline
        c->n_pix = (int)((W - 2 * (size_t)params->border) * (size_t)c->n_rows);
        c->tiles_x = (int)((W - 2 * (size_t)params->border + dmf::TILE_W - 1) / dmf::TILE_W);
        c->n_bands = (c->n_rows + dmf::TILE_H - 1) / dmf::TILE_H;
        if (c->n_bands < 1) c->n_bands = 1;
        c->n_ctas = c->tiles_x * c->n_bands;
        c->n_slots = c->n_ctas * dmf::TILE_PIX;  // >= n_pix (edge tiles are partly empty)
        // work units are (slot << CHUNK_BITS) | chunk in 32 bits: the slot index (not the pixel count) must fit 26 bits
        if ((long long)c->n_ctas * dmf::TILE_PIX > (1ll << (32 - dmf::CHUNK_BITS))) {
            int rc_ = fail(nullptr, DMF_ERR_INVALID, "dmf_create: the rows of this context need more than 2^26 pixel slots (split the image into more contexts)");
            dmf_destroy(ctx);
            return rc_;
        }
        const size_t np = (size_t)c->n_slots;
        const int n_max = (int)(2.0 * params->max_half_len / params->step) + 2;  // trip-count bound of ref:432
        const size_t max_full = (size_t)(n_max / dmf::CHUNK) + 1;
        for (int b = 0; b < 2; ++b) {
            CUX(cudaMalloc(&c->d_rec[b], np * sizeof(dmf::PixelRec)));
            CUX(cudaMalloc(&c->d_best[b], np * sizeof(unsigned long long)));
            CUX(cudaMalloc(&c->d_state_c[b], np * sizeof(double2)));
        }
        CUX(cudaMalloc(&c->d_units_full, np * max_full * sizeof(unsigned int)));
        CUX(cudaMalloc(&c->d_units_tail, np * (dmf::CHUNK - 1) * sizeof(unsigned int)));
        CUX(cudaMalloc(&c->d_ctrl, 3 * sizeof(dmf::Ctrl)));
        for (int b = 0; b < 2; ++b) {
            CUX(cudaMalloc(&c->d_mom1[b], W * H * sizeof(int4)));
            CUX(cudaMalloc(&c->d_mom2[b], W * H * sizeof(dmf::mom2_t)));
            CUX(cudaMalloc(&c->d_currx[b], W * H * sizeof(dmf::currx_t)));
            CUX(cudaMemsetAsync(c->d_mom1[b], 0, W * H * sizeof(int4), c->stream));
            CUX(cudaMemsetAsync(c->d_mom2[b], 0, W * H * sizeof(dmf::mom2_t), c->stream));
            CUX(cudaMemsetAsync(c->d_currx[b], 0, W * H * sizeof(dmf::currx_t), c->stream));
            CUX(cudaEventCreateWithFlags(&c->ev_mom_done[b], cudaEventDisableTiming));
            CUX(cudaEventCreateWithFlags(&c->ev_tab_free[b], cudaEventDisableTiming));
            CUX(cudaEventCreateWithFlags(&c->ev_adv_done[b], cudaEventDisableTiming));
        }
        CUX(cudaMalloc(&c->d_refx, W * H * sizeof(uint2)));
        CUX(cudaMemsetAsync(c->d_refx, 0, W * H * sizeof(uint2), c->stream));
        CUX(cudaMemsetAsync(c->d_ctrl, 0, 3 * sizeof(dmf::Ctrl), c->stream));
        CUX(cudaMalloc(&c->d_cta_cnt, 2 * (size_t)c->n_ctas * sizeof(unsigned int)));
        CUX(cudaMemsetAsync(c->d_cta_cnt, 0, 2 * (size_t)c->n_ctas * sizeof(unsigned int), c->stream));
        int per_sm = 0;
        // 4 CTAs / SM with the block shifted per sample (80 registers) where the per-frame tables exceed L2 (>= 4 M pixels:
        // measured -5 % on ncc_kernel at 4K), 3 CTAs / SM with the pre-shifted patch copy otherwise (DMF_NCC_SB=0/1 forces)
        const char *sbe = std::getenv("DMF_NCC_SB");
        const bool sb = sbe ? std::atoi(sbe) != 0 : (size_t)params->width * params->height >= (4u << 20);
        switch (params->width) {  // compile-time widths: the row loads of a sample become immediate offsets
            case 640: c->ncc_fn = sb ? dmf::ncc_kernel<640, true> : dmf::ncc_kernel<640, false>; break;
            case 1241: c->ncc_fn = sb ? dmf::ncc_kernel<1241, true> : dmf::ncc_kernel<1241, false>; break;
            case 1920: c->ncc_fn = sb ? dmf::ncc_kernel<1920, true> : dmf::ncc_kernel<1920, false>; break;
            case 3840: c->ncc_fn = sb ? dmf::ncc_kernel<3840, true> : dmf::ncc_kernel<3840, false>; break;
            default: c->ncc_fn = sb ? dmf::ncc_kernel<0, true> : dmf::ncc_kernel<0, false>; break;
        }
        {
            // Kernels that are to share an SM must agree on its L1 / shared-memory split: ncc_kernel uses no shared memory,
            // moments_bulk_kernel stages its tiles there; with different carve-outs the second kernel's CTAs wait until
            // the SM has drained (measured: no overlap at all).  Both ask for the same 32 KB-class carve-out.
            const char *cv = std::getenv("DMF_SMEM_CARVEOUT");
            const int carve = cv ? std::atoi(cv) : 14;
            if (carve >= 0) {
                CUX(cudaFuncSetAttribute(c->ncc_fn, cudaFuncAttributePreferredSharedMemoryCarveout, carve));
                CUX(cudaFuncSetAttribute(dmf::moments_bulk_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, carve));
            }
        }
        CUX(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, c->ncc_fn, dmf::NCC_THREADS, 0));
        if (per_sm < 1) per_sm = 1;
        c->ncc_grid = prop.multiProcessorCount * per_sm;  // persistent CTAs: one resident wave
        {
            const char *mm = std::getenv("DMF_MOMENTS"), *mc = std::getenv("DMF_MOMENTS_CTAS_PER_SM");
            // moments_bulk_kernel (tiles staged by bulk asynchronous copies) for frames with enough tiles to feed every SM
            // (>= 1 M pixels: 2 tiles per SM); the per-column kernel for small frames.  Both write the same bits; measured
            // equal at 1080p / 4K (profiles/r02_ab_ncc_prefetch_and_moments.txt), the per-column kernel 3 % ahead at 640x480.
            c->mom_bulk = mm ? std::strcmp(mm, "legacy") != 0 : (size_t)params->width * params->height >= (1u << 20);
            const char *mg = std::getenv("DMF_MOMENTS_GATE");
            c->mom_gate = !(mg && std::strcmp(mg, "0") == 0);
            if (const char *tp = std::getenv("DMF_TRACE")) {
                c->trace_path = tp;
                c->trace_cap = 1024;
                CUX(cudaMalloc(&c->d_trace, (size_t)c->trace_cap * 6 * sizeof(unsigned long long)));
                CUX(cudaMemset(c->d_trace, 0, (size_t)c->trace_cap * 6 * sizeof(unsigned long long)));
            }
            if (const char *mr = std::getenv("DMF_MOMENTS_REPEAT")) c->mom_repeat = std::atoi(mr) > 1 ? std::atoi(mr) : 1;
            int k = mc ? std::atoi(mc) : 2;
            if (k < 1) k = 1;
            c->mom_grid = prop.multiProcessorCount * k;
        }
    }
    CUX(cudaStreamSynchronize(c->stream));
#undef CUX
    *out = ctx;
    return DMF_OK;
}

int dmf_create(const dmf_params *params, int device, int row_begin, int row_end, dmf_ctx **out) {
    return create_common(params, device, row_begin, row_end, 1, 1, 0, out);
}

int dmf_create_cyclic(const dmf_params *params, int device, int block_rows, int n_parts, int part, dmf_ctx **out) {
    return create_common(params, device, 0, params ? params->height : 0, block_rows, n_parts, part, out);
}

int dmf_get_rows(const dmf_ctx *ctx, int *rows_out, int capacity, int *n_rows) {
    if (!ctx || !n_rows) return fail(nullptr, DMF_ERR_INVALID, "dmf_get_rows: NULL argument");
    int n = 0;
    for (const auto &sp : ctx->spans)
        for (int y = sp.first; y < sp.second; ++y, ++n)
            if (rows_out && n < capacity) rows_out[n] = y;
    *n_rows = n;
    return DMF_OK;
}

void dmf_destroy(dmf_ctx *ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    if (ctx->copy_stream) cudaStreamSynchronize(ctx->copy_stream);
    if (ctx->mom_stream) cudaStreamSynchronize(ctx->mom_stream);
    if (ctx->d_trace) {
        std::vector<unsigned long long> h((size_t)ctx->trace_cap * 6);
        if (cudaMemcpy(h.data(), ctx->d_trace, h.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost) == cudaSuccess)
            if (FILE *f = std::fopen(ctx->trace_path.c_str(), "w")) {
                std::fprintf(f, "# update advance_begin advance_end moments_eligible moments_end ncc_begin ncc_end   (ns, %%globaltimer)\n");
                for (int u = 0; u < ctx->trace_cap && h[6 * (size_t)u]; ++u)
                    std::fprintf(f, "%d %llu %llu %llu %llu %llu %llu\n", u, h[6 * u], h[6 * u + 1], h[6 * u + 2], h[6 * u + 3], h[6 * u + 4], h[6 * u + 5]);
                std::fclose(f);
            }
        cudaFree(ctx->d_trace);
    }
    cudaFree(ctx->d_ref);
    for (int b = 0; b < 2; ++b) {
        cudaFree(ctx->d_curr[b]);
        if (ctx->h_stage[b]) cudaFreeHost(ctx->h_stage[b]);
        if (ctx->ev_copied[b]) cudaEventDestroy(ctx->ev_copied[b]);
        if (ctx->ev_consumed[b]) cudaEventDestroy(ctx->ev_consumed[b]);
    }
    if (ctx->ev_ext) cudaEventDestroy(ctx->ev_ext);
    for (int b = 0; b < 2; ++b) if (ctx->h_shadow[b]) cudaFreeHost(ctx->h_shadow[b]);
    for (cudaEvent_t e : ctx->ev_pool) cudaEventDestroy(e);
    for (cudaEvent_t e : ctx->ev_flush) cudaEventDestroy(e);
    cudaFree(ctx->d_refstat); cudaFree(ctx->d_depth); cudaFree(ctx->d_cov2); cudaFree(ctx->d_truth);
    cudaFree(ctx->d_dbg_ncc); cudaFree(ctx->d_dbg_n);
    for (int b = 0; b < 2; ++b) { cudaFree(ctx->d_rec[b]); cudaFree(ctx->d_best[b]); cudaFree(ctx->d_state_c[b]); }
    cudaFree(ctx->d_units_full); cudaFree(ctx->d_units_tail); cudaFree(ctx->d_ctrl); cudaFree(ctx->d_cta_cnt); cudaFree(ctx->d_refx);
    for (int b = 0; b < 2; ++b) {
        cudaFree(ctx->d_mom1[b]); cudaFree(ctx->d_mom2[b]); cudaFree(ctx->d_currx[b]);
        if (ctx->ev_mom_done[b]) cudaEventDestroy(ctx->ev_mom_done[b]);
        if (ctx->ev_tab_free[b]) cudaEventDestroy(ctx->ev_tab_free[b]);
        if (ctx->ev_adv_done[b]) cudaEventDestroy(ctx->ev_adv_done[b]);
    }
    if (ctx->ev_frame) cudaEventDestroy(ctx->ev_frame);
    cudaFree(ctx->d_flags); cudaFree(ctx->d_mask); cudaFree(ctx->d_counters); cudaFree(ctx->d_eval);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
    if (ctx->mom_stream) cudaStreamDestroy(ctx->mom_stream);
    delete ctx;
}

int dmf_get_params(const dmf_ctx *ctx, dmf_params *out) {
    if (!ctx || !out) return fail(nullptr, DMF_ERR_INVALID, "dmf_get_params: NULL argument");
    *out = ctx->prm;
    return DMF_OK;
}

int dmf_get_band(const dmf_ctx *ctx, int *row_begin, int *row_end) {
    if (!ctx || !row_begin || !row_end) return fail(nullptr, DMF_ERR_INVALID, "dmf_get_band: NULL argument");
    *row_begin = ctx->spans.empty() ? ctx->row0 : ctx->spans.front().first;
    *row_end = ctx->spans.empty() ? ctx->row0 : ctx->spans.back().second;
    return DMF_OK;
}

static int run_ref_stats(dmf_ctx *c) {
    const dmf_params &p = c->prm;
    dim3 blk(32, 8);
    dim3 grid((p.width - 2 * p.border + 31) / 32, (p.height - 2 * p.border + 7) / 8);
    dmf::ref_stats_kernel<<<grid, blk, 0, c->stream>>>(c->d_ref, c->img_pitch, p.width, p.height, p.border, c->d_refstat, p.width);
    dmf::ref_expand_kernel<<<dim3((p.width + 255) / 256, p.height), 256, 0, c->stream>>>(c->d_ref, c->img_pitch, p.width, p.height, c->d_refx);
    CU(cudaGetLastError());
    c->have_ref = true;
    return DMF_OK;
}

int dmf_set_reference(dmf_ctx *c, const uint8_t *ref_host, size_t step) {
    if (!c || !ref_host) return fail(c, DMF_ERR_INVALID, "dmf_set_reference: NULL argument");
    if (step < (size_t)c->prm.width) return fail(c, DMF_ERR_INVALID, "dmf_set_reference: step < width");
    CU(cudaSetDevice(c->device));
    { int rc_ = flush_pending(c); if (rc_) return rc_; }
    CU(cudaMemcpy2DAsync(c->d_ref, c->img_pitch, ref_host, step, c->prm.width, c->prm.height, cudaMemcpyHostToDevice, c->stream));
    CU(cudaStreamSynchronize(c->stream));  // the host image may be released on return
    c->ref_hash_valid = false;
    return run_ref_stats(c);
}

int dmf_set_reference_device(dmf_ctx *c, const uint8_t *ref_dev, size_t step) {
    if (!c || !ref_dev) return fail(c, DMF_ERR_INVALID, "dmf_set_reference_device: NULL argument");
    if (step < (size_t)c->prm.width) return fail(c, DMF_ERR_INVALID, "dmf_set_reference_device: step < width");
    CU(cudaSetDevice(c->device));
    { int rc_ = flush_pending(c); if (rc_) return rc_; }
    CU(cudaMemcpy2DAsync(c->d_ref, c->img_pitch, ref_dev, step, c->prm.width, c->prm.height, cudaMemcpyDeviceToDevice, c->stream));
    c->ref_hash_valid = false;
    return run_ref_stats(c);
}

int dmf_fill_state(dmf_ctx *c, double init_depth, double init_cov2) {
    if (!c) return fail(c, DMF_ERR_INVALID, "dmf_fill_state: NULL context");
    CU(cudaSetDevice(c->device));
    { int rc_ = flush_pending(c); if (rc_) return rc_; }
    const size_t n = (size_t)c->prm.width * c->prm.height;
    c->shadow_valid = false;
    dmf::fill_state_kernel<<<148 * 4, 256, 0, c->stream>>>(c->d_depth, c->d_cov2, n, init_depth, init_cov2);
    CU(cudaGetLastError());
    return DMF_OK;
}

int dmf_upload_state(dmf_ctx *c, const double *depth, size_t depth_step, const double *cov2, size_t cov2_step) {
    if (!c || !depth || !cov2) return fail(c, DMF_ERR_INVALID, "dmf_upload_state: NULL argument");
    const size_t rowb = (size_t)c->prm.width * sizeof(double);
    if (depth_step < rowb || cov2_step < rowb) return fail(c, DMF_ERR_INVALID, "dmf_upload_state: step < width*8");
    CU(cudaSetDevice(c->device));
    { int rc_ = flush_pending(c); if (rc_) return rc_; }
    c->shadow_valid = false;
    for (const auto &sp : c->io_spans) {
        const int y0 = sp.first, rows = sp.second - sp.first;
        CU(cudaMemcpy2DAsync(c->d_depth + (size_t)y0 * c->prm.width, rowb, (const char *)depth + (size_t)y0 * depth_step, depth_step, rowb, rows, cudaMemcpyHostToDevice, c->stream));
        CU(cudaMemcpy2DAsync(c->d_cov2 + (size_t)y0 * c->prm.width, rowb, (const char *)cov2 + (size_t)y0 * cov2_step, cov2_step, rowb, rows, cudaMemcpyHostToDevice, c->stream));
    }
    CU(cudaStreamSynchronize(c->stream));
    return DMF_OK;
}

int dmf_download_state(dmf_ctx *c, double *depth, size_t depth_step, double *cov2, size_t cov2_step) {
    if (!c || !depth || !cov2) return fail(c, DMF_ERR_INVALID, "dmf_download_state: NULL argument");
    const size_t rowb = (size_t)c->prm.width * sizeof(double);
    if (depth_step < rowb || cov2_step < rowb) return fail(c, DMF_ERR_INVALID, "dmf_download_state: step < width*8");
    CU(cudaSetDevice(c->device));
    { int rc_ = flush_pending(c); if (rc_) return rc_; }
    for (const auto &sp : c->io_spans) {
        const int y0 = sp.first, rows = sp.second - sp.first;
        CU(cudaMemcpy2DAsync((char *)depth + (size_t)y0 * depth_step, depth_step, c->d_depth + (size_t)y0 * c->prm.width, rowb, rowb, rows, cudaMemcpyDeviceToHost, c->stream));
        CU(cudaMemcpy2DAsync((char *)cov2 + (size_t)y0 * cov2_step, cov2_step, c->d_cov2 + (size_t)y0 * c->prm.width, rowb, rowb, rows, cudaMemcpyDeviceToHost, c->stream));
    }
    CU(cudaStreamSynchronize(c->stream));
    return DMF_OK;
}

int dmf_update(dmf_ctx *c, const uint8_t *curr_host, size_t step, const double q[4], const double t[3]) {
    if (!c || !curr_host || !q || !t) return fail(c, DMF_ERR_INVALID, "dmf_update: NULL argument");
    if (!c->have_ref) return fail(c, DMF_ERR_STATE, "dmf_update: dmf_set_reference() has not been called");
    if (step < (size_t)c->prm.width) return fail(c, DMF_ERR_INVALID, "dmf_update: step < width");
    CU(cudaSetDevice(c->device));
    const int b = (int)(c->frame_idx & 1);
    c->frame_idx++;
    const int W = c->prm.width, H = c->prm.height;
    // Is the caller's frame pinned (dmf_alloc_pinned / cudaHostRegister)?  Then copy straight from it.
    cudaPointerAttributes attr{};
    bool pinned = false;
    if (cudaPointerGetAttributes(&attr, curr_host) == cudaSuccess) pinned = (attr.type == cudaMemoryTypeHost);
    else cudaGetLastError();
    const uint8_t *src = curr_host;
    size_t src_step = step;
    if (!pinned) {
        // staging buffer b is free once its previous H2D copy has completed
        CU(cudaEventSynchronize(c->ev_copied[b]));
        if (step == (size_t)c->img_pitch) std::memcpy(c->h_stage[b], curr_host, (size_t)step * H);
        else for (int y = 0; y < H; ++y) std::memcpy(c->h_stage[b] + (size_t)y * c->img_pitch, curr_host + (size_t)y * step, W);
        src = c->h_stage[b];
        src_step = c->img_pitch;
    }
    // the device buffer b is free once the kernel of two frames ago has consumed it
    CU(cudaStreamWaitEvent(c->copy_stream, c->ev_consumed[b], 0));
    if (src_step == (size_t)c->img_pitch && (size_t)W == src_step)  // contiguous frame: one linear DMA instead of H row descriptors
        CU(cudaMemcpyAsync(c->d_curr[b], src, src_step * (size_t)H, cudaMemcpyHostToDevice, c->copy_stream));
    else
        CU(cudaMemcpy2DAsync(c->d_curr[b], c->img_pitch, src, src_step, W, H, cudaMemcpyHostToDevice, c->copy_stream));
    CU(cudaEventRecord(c->ev_copied[b], c->copy_stream));
    return launch_update(c, c->d_curr[b], c->img_pitch, q, t, c->ev_copied[b], c->ev_consumed[b]);
}

int dmf_update_device(dmf_ctx *c, const uint8_t *curr_dev, size_t step, const double q[4], const double t[3], void *wait_stream) {
    if (!c || !curr_dev || !q || !t) return fail(c, DMF_ERR_INVALID, "dmf_update_device: NULL argument");
    if (!c->have_ref) return fail(c, DMF_ERR_STATE, "dmf_update_device: dmf_set_reference() has not been called");
    if (step < (size_t)c->prm.width) return fail(c, DMF_ERR_INVALID, "dmf_update_device: step < width");
    CU(cudaSetDevice(c->device));
    cudaEvent_t ready = nullptr;
    if (wait_stream) {
        CU(cudaEventRecord(c->ev_ext, (cudaStream_t)wait_stream));
        ready = c->ev_ext;
    }
    const bool aligned = ((reinterpret_cast<uintptr_t>(curr_dev) & 3u) == 0) && (step % 4 == 0) && step <= 0x7fffffff;
    if (aligned) return launch_update(c, curr_dev, (int)step, q, t, ready, nullptr);
    // unaligned device frame: repack into an internal pitched buffer on the copy stream
    const int b = (int)(c->frame_idx & 1);
    c->frame_idx++;
    if (ready) CU(cudaStreamWaitEvent(c->copy_stream, ready, 0));
    CU(cudaStreamWaitEvent(c->copy_stream, c->ev_consumed[b], 0));
    CU(cudaMemcpy2DAsync(c->d_curr[b], c->img_pitch, curr_dev, step, c->prm.width, c->prm.height, cudaMemcpyDeviceToDevice, c->copy_stream));
    CU(cudaEventRecord(c->ev_copied[b], c->copy_stream));
    return launch_update(c, c->d_curr[b], c->img_pitch, q, t, c->ev_copied[b], c->ev_consumed[b]);
}

// 64-bit content hash of an image (4 interleaved FNV-1a style lanes over 8-byte words; ~10 GB/s on one core)
static unsigned long long image_hash(const uint8_t *img, size_t step, int width, int height) {
    unsigned long long h[4] = {0xcbf29ce484222325ull, 0x9e3779b97f4a7c15ull, 0xc2b2ae3d27d4eb4full, 0x165667b19e3779f9ull};
    for (int y = 0; y < height; ++y) {
        const uint8_t *row = img + (size_t)y * step;
        int x = 0;
        for (; x + 32 <= width; x += 32)
            for (int k = 0; k < 4; ++k) {
                unsigned long long w;
                std::memcpy(&w, row + x + 8 * k, 8);
                h[k] = (h[k] ^ w) * 0x100000001b3ull;
                h[k] ^= h[k] >> 29;
            }
        for (; x < width; ++x) h[0] = (h[0] ^ row[x]) * 0x100000001b3ull;
    }
    return (h[0] ^ (h[1] << 1) ^ (h[2] << 2) ^ (h[3] << 3)) + (unsigned long long)width * 1315423911ull + (unsigned long long)height;
}

#ifndef STRICT_HOST_THREADS
#define STRICT_HOST_THREADS 4
#endif
int dmf_update_strict(dmf_ctx *c, const uint8_t *ref_host, size_t ref_step, const uint8_t *curr_host, size_t curr_step,
                      const double q[4], const double t[3], double *depth, size_t depth_step, double *cov2, size_t cov2_step) {
    if (!c || !ref_host || !curr_host || !q || !t || !depth || !cov2) return fail(c, DMF_ERR_INVALID, "dmf_update_strict: NULL argument");
    const int W = c->prm.width, H = c->prm.height;
    const size_t rowb = (size_t)W * sizeof(double);
    if (ref_step < (size_t)W || curr_step < (size_t)W || depth_step < rowb || cov2_step < rowb)
        return fail(c, DMF_ERR_INVALID, "dmf_update_strict: a step is smaller than its row");
    if (c->n_rows != H - 2 * c->prm.border) return fail(c, DMF_ERR_STATE, "dmf_update_strict: needs a context that owns every interior row");
    CU(cudaSetDevice(c->device));
    // 1. the reference image: uploaded (and its patch statistics recomputed) only when its content changed
    const unsigned long long hr = image_hash(ref_host, ref_step, W, H);
    if (!c->have_ref || !c->ref_hash_valid || hr != c->ref_hash) {
        int rc = dmf_set_reference(c, ref_host, ref_step);
        if (rc) return rc;
        c->ref_hash = hr;
        c->ref_hash_valid = true;
    }
    // 2. the maps: the caller may have modified them between calls (they are plain in/out arguments, ref:107-112);
    //    rows that still equal what the previous call returned are already in HBM
    for (int b = 0; b < 2; ++b)
        if (!c->h_shadow[b]) { CU(cudaMallocHost(&c->h_shadow[b], rowb * H)); c->shadow_valid = false; }
    // (the row loops of this function run on a few host threads: they move 16*W*H bytes per call and were, single-threaded,
    //  the larger half of a 640x480 call)
    int same = c->shadow_valid ? 1 : 0;
    if (same) {
#pragma omp parallel for num_threads(STRICT_HOST_THREADS) reduction(&& : same) schedule(static)
        for (int y = 0; y < H; ++y)
            same = same && std::memcmp((const char *)depth + (size_t)y * depth_step, c->h_shadow[0] + (size_t)y * W, rowb) == 0 &&
                   std::memcmp((const char *)cov2 + (size_t)y * cov2_step, c->h_shadow[1] + (size_t)y * W, rowb) == 0;
    }
    if (!same) {
        { int rc_ = flush_pending(c); if (rc_) return rc_; }
#pragma omp parallel for num_threads(STRICT_HOST_THREADS) schedule(static)
        for (int y = 0; y < H; ++y) {
            std::memcpy(c->h_shadow[0] + (size_t)y * W, (const char *)depth + (size_t)y * depth_step, rowb);
            std::memcpy(c->h_shadow[1] + (size_t)y * W, (const char *)cov2 + (size_t)y * cov2_step, rowb);
        }
        CU(cudaMemcpyAsync(c->d_depth, c->h_shadow[0], rowb * H, cudaMemcpyHostToDevice, c->stream));
        CU(cudaMemcpyAsync(c->d_cov2, c->h_shadow[1], rowb * H, cudaMemcpyHostToDevice, c->stream));
    }
    // 3. the update itself
    { int rc = dmf_update(c, curr_host, curr_step, q, t); if (rc) return rc; }
    // 4. both maps valid in the caller's memory on return (ref:292-300 reads them after every call): the cov2 download
    //    overlaps the host copy of depth
    { int rc_ = flush_pending(c); if (rc_) return rc_; }
    CU(cudaMemcpyAsync(c->h_shadow[0], c->d_depth, rowb * H, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaEventRecord(c->ev_frame, c->stream));
    CU(cudaMemcpyAsync(c->h_shadow[1], c->d_cov2, rowb * H, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaEventSynchronize(c->ev_frame));
#pragma omp parallel for num_threads(STRICT_HOST_THREADS) schedule(static)
    for (int y = 0; y < H; ++y) std::memcpy((char *)depth + (size_t)y * depth_step, c->h_shadow[0] + (size_t)y * W, rowb);
    CU(cudaStreamSynchronize(c->stream));
#pragma omp parallel for num_threads(STRICT_HOST_THREADS) schedule(static)
    for (int y = 0; y < H; ++y) std::memcpy((char *)cov2 + (size_t)y * cov2_step, c->h_shadow[1] + (size_t)y * W, rowb);
    c->shadow_valid = true;
    return DMF_OK;
}

int dmf_flush(dmf_ctx *c) {
    if (!c) return fail(c, DMF_ERR_INVALID, "dmf_flush: NULL context");
    return flush_pending(c);
}

int dmf_sync(dmf_ctx *c) {
    if (!c) return fail(c, DMF_ERR_INVALID, "dmf_sync: NULL context");
    CU(cudaSetDevice(c->device));
    { int rc_ = flush_pending(c); if (rc_) return rc_; }
    CU(cudaStreamSynchronize(c->copy_stream));
    CU(cudaStreamSynchronize(c->mom_stream));
    CU(cudaStreamSynchronize(c->stream));
    return DMF_OK;
}

int dmf_read_counters(dmf_ctx *c, dmf_counters *out, int reset) {
    if (!c || !out) return fail(c, DMF_ERR_INVALID, "dmf_read_counters: NULL argument");
    CU(cudaSetDevice(c->device));
    { int rc_ = flush_pending(c); if (rc_) return rc_; }
    unsigned long long h[3];
    CU(cudaMemcpyAsync(h, c->d_counters, sizeof(h), cudaMemcpyDeviceToHost, c->stream));
    if (reset) CU(cudaMemsetAsync(c->d_counters, 0, sizeof(h), c->stream));
    CU(cudaStreamSynchronize(c->stream));
    out->frames = c->frames;
    out->interior = c->frames * (unsigned long long)c->n_rows * (unsigned long long)(c->prm.width - 2 * c->prm.border);
    out->active = h[0]; out->ncc_evals = h[1]; out->accepted = h[2];
    if (reset) c->frames = 0;
    return DMF_OK;
}

int dmf_set_timing(dmf_ctx *c, int enable) {
    if (!c) return fail(c, DMF_ERR_INVALID, "dmf_set_timing: NULL context");
    c->timing_on = enable != 0;
    return DMF_OK;
}

int dmf_get_timing(dmf_ctx *c, double ms_out[4], uint64_t *frames, int reset) {
    if (!c || !ms_out || !frames) return fail(c, DMF_ERR_INVALID, "dmf_get_timing: NULL argument");
    CU(cudaSetDevice(c->device));
    CU(cudaStreamSynchronize(c->stream));
    for (size_t i = 0; i + 4 < c->ev_used + 1 && i + 4 < c->ev_pool.size() + 1 && i + 5 <= c->ev_used; i += 5) {
        for (int k = 0; k < 4; ++k) {
            float ms = 0.f;
            CU(cudaEventElapsedTime(&ms, c->ev_pool[i + k], c->ev_pool[i + k + 1]));
            c->timing_ms[k] += ms;
        }
        c->timing_frames++;
    }
    c->ev_used = 0;
    for (size_t i = 0; i + 1 < c->ev_flush_used; i += 2) {
        float ms = 0.f;
        CU(cudaEventElapsedTime(&ms, c->ev_flush[i], c->ev_flush[i + 1]));
        c->timing_ms[3] += ms;
    }
    c->ev_flush_used = 0;
    for (int k = 0; k < 4; ++k) ms_out[k] = c->timing_ms[k];
    *frames = c->timing_frames;
    if (reset) { for (int k = 0; k < 4; ++k) c->timing_ms[k] = 0; c->timing_frames = 0; }
    return DMF_OK;
}

int dmf_enable_flags(dmf_ctx *c, int enable) {
    if (!c) return fail(c, DMF_ERR_INVALID, "dmf_enable_flags: NULL context");
    CU(cudaSetDevice(c->device));
    { int rc_ = flush_pending(c); if (rc_) return rc_; }
    if (enable && !c->d_dbg_ncc) {
        const size_t n = (size_t)c->prm.width * c->prm.height;
        CU(cudaMalloc(&c->d_dbg_ncc, n * sizeof(float)));
        CU(cudaMalloc(&c->d_dbg_n, n * sizeof(int)));
        CU(cudaMemsetAsync(c->d_dbg_ncc, 0, n * sizeof(float), c->stream));
        CU(cudaMemsetAsync(c->d_dbg_n, 0, n * sizeof(int), c->stream));
    }
    c->flags_on = enable != 0;
    return DMF_OK;
}

int dmf_download_flags(dmf_ctx *c, uint8_t *flags_host, size_t step) {
    if (!c || !flags_host) return fail(c, DMF_ERR_INVALID, "dmf_download_flags: NULL argument");
    if (step < (size_t)c->prm.width) return fail(c, DMF_ERR_INVALID, "dmf_download_flags: step < width");
    CU(cudaSetDevice(c->device));
    { int rc_ = flush_pending(c); if (rc_) return rc_; }
    for (const auto &sp : c->io_spans) {
        const int y0 = sp.first, rows = sp.second - sp.first;
        CU(cudaMemcpy2DAsync(flags_host + (size_t)y0 * step, step, c->d_flags + (size_t)y0 * c->prm.width, c->prm.width, c->prm.width, rows, cudaMemcpyDeviceToHost, c->stream));
    }
    CU(cudaStreamSynchronize(c->stream));
    return DMF_OK;
}

int dmf_download_debug(dmf_ctx *c, float *best_ncc_host, int32_t *samples_host) {
    if (!c || !best_ncc_host || !samples_host) return fail(c, DMF_ERR_INVALID, "dmf_download_debug: NULL argument");
    if (!c->d_dbg_ncc) return fail(c, DMF_ERR_STATE, "dmf_download_debug: dmf_enable_flags(ctx, 1) has not been called");
    CU(cudaSetDevice(c->device));
    { int rc_ = flush_pending(c); if (rc_) return rc_; }
    const size_t n = (size_t)c->prm.width * c->prm.height;
    CU(cudaMemcpyAsync(best_ncc_host, c->d_dbg_ncc, n * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaMemcpyAsync(samples_host, c->d_dbg_n, n * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return DMF_OK;
}

int dmf_device_state(dmf_ctx *c, double **depth_dev, double **cov2_dev, size_t *pitch) {
    if (!c || !depth_dev || !cov2_dev || !pitch) return fail(c, DMF_ERR_INVALID, "dmf_device_state: NULL argument");
    { int rc_ = flush_pending(c); if (rc_) return rc_; }
    *depth_dev = c->d_depth; *cov2_dev = c->d_cov2; *pitch = (size_t)c->prm.width * sizeof(double);
    return DMF_OK;
}

int dmf_stream(dmf_ctx *c, void **stream) {
    if (!c || !stream) return fail(c, DMF_ERR_INVALID, "dmf_stream: NULL argument");
    *stream = (void *)c->stream;
    return DMF_OK;
}

int dmf_selftest_division(int device, uint64_t n, uint64_t seed, uint64_t *mismatches) {
    dmf_ctx_impl *c = nullptr;
    if (!mismatches) return fail(c, DMF_ERR_INVALID, "dmf_selftest_division: NULL argument");
    CU(cudaSetDevice(device));
    unsigned long long *d = nullptr, h = 0;
    CU(cudaMalloc(&d, sizeof(h)));
    CU(cudaMemset(d, 0, sizeof(h)));
    dmf::division_selftest_kernel<<<148 * 8, 256>>>((unsigned long long)seed, (unsigned long long)n, d);
    cudaError_t e = cudaMemcpy(&h, d, sizeof(h), cudaMemcpyDeviceToHost);
    cudaFree(d);
    if (e != cudaSuccess) return fail(c, DMF_ERR_CUDA, std::string("dmf_selftest_division: ") + cudaGetErrorString(e));
    *mismatches = h;
    return DMF_OK;
}

int dmf_alloc_pinned(void **ptr, size_t bytes) {
    dmf_ctx_impl *c = nullptr;
    if (!ptr) return fail(c, DMF_ERR_INVALID, "dmf_alloc_pinned: NULL argument");
    CU(cudaMallocHost(ptr, bytes));
    return DMF_OK;
}

int dmf_host_register(void *ptr, size_t bytes) {
    dmf_ctx_impl *c = nullptr;
    if (!ptr || !bytes) return fail(c, DMF_ERR_INVALID, "dmf_host_register: NULL argument");
    CU(cudaHostRegister(ptr, bytes, cudaHostRegisterPortable));
    return DMF_OK;
}

int dmf_host_unregister(void *ptr) {
    dmf_ctx_impl *c = nullptr;
    if (ptr) CU(cudaHostUnregister(ptr));
    return DMF_OK;
}

int dmf_free_pinned(void *ptr) {
    dmf_ctx_impl *c = nullptr;
    if (ptr) CU(cudaFreeHost(ptr));
    return DMF_OK;
}

int dmf_set_truth(dmf_ctx *c, const double *truth_host, size_t step) {
    if (!c || !truth_host) return fail(c, DMF_ERR_INVALID, "dmf_set_truth: NULL argument");
    const size_t rowb = (size_t)c->prm.width * sizeof(double);
    if (step < rowb) return fail(c, DMF_ERR_INVALID, "dmf_set_truth: step < width*8");
    CU(cudaSetDevice(c->device));
    if (!c->d_truth) CU(cudaMalloc(&c->d_truth, rowb * c->prm.height));
    CU(cudaMemcpy2DAsync(c->d_truth, rowb, truth_host, step, rowb, c->prm.height, cudaMemcpyHostToDevice, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    c->have_truth = true;
    return DMF_OK;
}

int dmf_evaluate_depth(dmf_ctx *c, double max_variance, double *sum_sq, uint64_t *count) {
    if (!c || !sum_sq || !count) return fail(c, DMF_ERR_INVALID, "dmf_evaluate_depth: NULL argument");
    if (!c->have_truth) return fail(c, DMF_ERR_STATE, "dmf_evaluate_depth: dmf_set_truth() has not been called");
    CU(cudaSetDevice(c->device));
    { int rc_ = flush_pending(c); if (rc_) return rc_; }
    CU(cudaMemsetAsync(c->d_eval, 0, sizeof(double), c->stream));
    CU(cudaMemsetAsync(c->d_counters + 3, 0, sizeof(unsigned long long), c->stream));
    for (const auto &sp : c->spans) {
        dmf::evaluate_depth_kernel<<<148 * 2, 256, 0, c->stream>>>(c->d_truth, c->d_depth, c->d_cov2, c->prm.width, c->prm.border,
                                                                   c->prm.width - c->prm.border, sp.first, sp.second, max_variance,
                                                                   c->d_eval, c->d_counters + 3);
        CU(cudaGetLastError());
    }
    unsigned long long n = 0;
    CU(cudaMemcpyAsync(sum_sq, c->d_eval, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaMemcpyAsync(&n, c->d_counters + 3, sizeof(n), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    *count = n;
    return DMF_OK;
}

int dmf_variance_mask(dmf_ctx *c, double max_variance, uint8_t *mask_host, size_t step) {
    if (!c || !mask_host) return fail(c, DMF_ERR_INVALID, "dmf_variance_mask: NULL argument");
    if (step < (size_t)c->prm.width) return fail(c, DMF_ERR_INVALID, "dmf_variance_mask: step < width");
    CU(cudaSetDevice(c->device));
    { int rc_ = flush_pending(c); if (rc_) return rc_; }
    const int W = c->prm.width;
    if (!c->d_mask) CU(cudaMalloc(&c->d_mask, (size_t)W * c->prm.height));
    for (const auto &sp : c->io_spans) {
        const int y0 = sp.first, y1 = sp.second;
        dim3 blk(64, 4), grid((W + 63) / 64, (y1 - y0 + 3) / 4);
        dmf::variance_mask_kernel<<<grid, blk, 0, c->stream>>>(c->d_cov2, W, W, y0, y1, max_variance, c->d_mask, W);
        CU(cudaGetLastError());
        CU(cudaMemcpy2DAsync(mask_host + (size_t)y0 * step, step, c->d_mask + (size_t)y0 * W, W, W, y1 - y0, cudaMemcpyDeviceToHost, c->stream));
    }
    CU(cudaStreamSynchronize(c->stream));
    return DMF_OK;
}

int dmf_point_cloud(dmf_ctx *c, const uint8_t *color_host, size_t color_step, int channels, double max_variance,
                    float *xyz_host, uint8_t *rgb_host, uint64_t capacity, uint64_t *n_points) {
    if (!c || !color_host || !xyz_host || !rgb_host || !n_points) return fail(c, DMF_ERR_INVALID, "dmf_point_cloud: NULL argument");
    if (channels != 1 && channels != 3 && channels != 4) return fail(c, DMF_ERR_INVALID, "dmf_point_cloud: channels must be 1, 3 or 4");
    const dmf_params &p = c->prm;
    if (color_step < (size_t)p.width * channels) return fail(c, DMF_ERR_INVALID, "dmf_point_cloud: step < width*channels");
    CU(cudaSetDevice(c->device));
    { int rc_ = flush_pending(c); if (rc_) return rc_; }
    const int n_rows = c->n_rows;
    *n_points = 0;
    if (n_rows <= 0) return DMF_OK;
    std::vector<int> rowlist;  // owned image rows, ascending: the reference's scan order restricted to this context
    for (const auto &sp : c->spans)
        for (int y = sp.first; y < sp.second; ++y) rowlist.push_back(y);
    uint8_t *d_color = nullptr, *d_rgb = nullptr;
    float *d_xyz = nullptr;
    unsigned int *d_rows = nullptr;
    int *d_rowlist = nullptr;
    const size_t cpitch = (size_t)p.width * channels;
    const uint64_t max_pts = (uint64_t)n_rows * (uint64_t)(p.width - 2 * p.border);
    const uint64_t cap = capacity < max_pts ? capacity : max_pts;
    int rc = DMF_OK;
    auto cleanup = [&]() { cudaFree(d_color); cudaFree(d_rgb); cudaFree(d_xyz); cudaFree(d_rows); cudaFree(d_rowlist); };
#define CUP(call)                                                                                          \
    do {                                                                                                   \
        cudaError_t e_ = (call);                                                                           \
        if (e_ != cudaSuccess) { rc = fail(c, DMF_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); cleanup(); return rc; } \
    } while (0)
    CUP(cudaMalloc(&d_color, cpitch * p.height));
    CUP(cudaMalloc(&d_rows, (size_t)(n_rows + 1) * sizeof(unsigned int)));
    CUP(cudaMalloc(&d_xyz, (cap ? cap : 1) * 3 * sizeof(float)));
    CUP(cudaMalloc(&d_rgb, (cap ? cap : 1) * 3));
    CUP(cudaMalloc(&d_rowlist, (size_t)n_rows * sizeof(int)));
    CUP(cudaMemcpyAsync(d_rowlist, rowlist.data(), (size_t)n_rows * sizeof(int), cudaMemcpyHostToDevice, c->stream));
    CUP(cudaMemcpy2DAsync(d_color, cpitch, color_host, color_step, cpitch, p.height, cudaMemcpyHostToDevice, c->stream));
    dmf::cloud_count_kernel<<<n_rows, 256, 0, c->stream>>>(c->d_depth, c->d_cov2, p.width, p.border, p.width - p.border, d_rowlist,
                                                           max_variance, d_rows);
    dmf::cloud_scan_kernel<<<1, 1024, 0, c->stream>>>(d_rows, n_rows);
    dmf::cloud_write_kernel<<<n_rows, 256, 0, c->stream>>>(c->d_depth, c->d_cov2, p.width, d_color, (int)cpitch, channels, p.border,
                                                           p.width - p.border, d_rowlist, max_variance, p.cx, p.cy, p.fx, p.fy, d_rows,
                                                           d_xyz, d_rgb, cap);
    CUP(cudaGetLastError());
    unsigned int total = 0;
    CUP(cudaMemcpyAsync(&total, d_rows + n_rows, sizeof(total), cudaMemcpyDeviceToHost, c->stream));
    CUP(cudaStreamSynchronize(c->stream));
    const uint64_t n_out = total < cap ? total : cap;
    if (n_out) {
        CUP(cudaMemcpyAsync(xyz_host, d_xyz, n_out * 3 * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
        CUP(cudaMemcpyAsync(rgb_host, d_rgb, n_out * 3, cudaMemcpyDeviceToHost, c->stream));
        CUP(cudaStreamSynchronize(c->stream));
    }
#undef CUP
    cleanup();
    *n_points = total;
    return DMF_OK;
}

}  // extern "C"
