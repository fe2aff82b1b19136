/*
 * dmf.h — C ABI of the B200-native dense monocular depth filter ("dmf").
 *
 * This is the drop-in boundary for ONE path of luigifreda/slamplay: the per-pixel
 * `update()` loop of dense_mapping/test_monocular_mapping.cpp.  The reference has no
 * FFI/plugin layer for this path; its boundary is the free function
 *
 *     void update(const Mat &ref, const Mat &curr, const SE3d &T_C_R,
 *                 Mat &depth, Mat &depth_cov2);
 *     (declared dense_mapping/test_monocular_mapping.cpp:107-112, defined :355-393,
 *      single call site :291)
 *
 * Every entry point below names the reference lines it replaces.  The C++ shim
 * `slamplay_b200/cpp/dense_mono_update.hpp` re-exposes exactly the reference
 * signature on cv::Mat / Sophus::SE3d (or layout-compatible stand-ins) and forwards
 * here; `slamplay_b200/depth_filter.py` is the ctypes binding used by tests/bench.
 *
 * Plain pointers and sizes only; no torch / OpenCV / Eigen types cross this ABI.
 * All functions return 0 on success or a negative dmf_status; the message is
 * available from dmf_last_error().  There is NO CPU fallback: every compute entry
 * point fails with DMF_ERR_CUDA when no sm_100 device is usable.
 */
#ifndef DMF_H_
#define DMF_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DMF_ABI_VERSION 1

typedef enum dmf_status {
    DMF_OK = 0,
    DMF_ERR_INVALID = -1,   /* bad argument (null pointer, size mismatch, bad params) */
    DMF_ERR_CUDA = -2,      /* CUDA runtime error / no usable device */
    DMF_ERR_STATE = -3,     /* call order violation (e.g. update before set_reference) */
    DMF_ERR_NOMEM = -4
} dmf_status;

/*
 * Runtime form of the file-scope constants of the reference
 * (dense_mapping/test_monocular_mapping.cpp:72-89) and of the literals used inside
 * epipolarSearch (:412-414 n_sigma/min_depth, :422 max_half_len, :432 step,
 * :443 ncc_thresh).  dmf_default_params() fills in the reference values.
 */
typedef struct dmf_params {
    int32_t width;          /* :73  */
    int32_t height;         /* :74  */
    int32_t border;         /* :72  (20); >= 13 */
    int32_t ncc_half;       /* :79  ncc_window_size (3); only 3 is supported by the kernels */
    double fx, fy, cx, cy;  /* :75-78 (float literals widened to double) */
    double step;            /* :432 0.7 */
    double max_half_len;    /* :422 100 */
    double min_depth;       /* :414 0.1 */
    double n_sigma;         /* :412 3 */
    double ncc_thresh;      /* :443 (double)0.85f */
    double min_cov;         /* :86  0.01*0.01   (inverse-depth variant :82  1e-4) */
    double max_cov;         /* :87  10          (inverse-depth variant :83  1)    */
    int32_t inverse_depth;  /* :63  USE_INVERSE_DEPTH_FOR_FILTERING (0 in the reference build) */
    int32_t reserved;
} dmf_params;

/* Work counters: they define the algorithmic work of a run. */
typedef struct dmf_counters {
    uint64_t frames;      /* update() calls since last reset */
    uint64_t interior;    /* pixels visited by the loops :357,:363 (band-local) */
    uint64_t active;      /* pixels passing the gate :366 */
    uint64_t ncc_evals;   /* NCC() calls :437 */
    uint64_t accepted;    /* epipolarSearch() returning true :443-446 */
} dmf_counters;

typedef struct dmf_ctx dmf_ctx;

/* Library/ABI version and build info ("sm_100a", compile date). Never fails. */
int dmf_abi_version(void);
const char *dmf_build_info(void);

/* Last error message of `ctx`, or of the calling thread when ctx == NULL. */
const char *dmf_last_error(const dmf_ctx *ctx);

/*
 * Reference constants :72-89 for a width x height image.  For 640x480 the
 * intrinsics are exactly the reference's float literals (fx = (double)481.2f, ...);
 * for other sizes they are scaled by width/640 with the principal point at the image
 * centre (SURVEY.md §8d synthetic inputs).  inverse_depth selects the :82-83 thresholds.
 */
int dmf_default_params(dmf_params *p, int width, int height, int inverse_depth);

/*
 * Create a filter context on CUDA device `device` that owns interior rows
 * [row_begin, row_end) of the depth / depth_cov2 maps (clamped to
 * [border, height-border), loop bounds of :357).  Pass 0, height for a single-GPU
 * context.  Allocates the state maps (:277-278) in HBM; they stay resident until
 * dmf_destroy().
 */
int dmf_create(const dmf_params *params, int device, int row_begin, int row_end, dmf_ctx **out);
/*
 * Block-cyclic row ownership for multi-GPU runs (SURVEY.md 8e "fallback if contiguous bands do not
 * balance"): the interior rows are cut into blocks of `block_rows` rows, dealt to `n_parts`
 * contexts in boustrophedon order (0..n-1, n-1..0, ...); this context is number `part`.  Convergence varies smoothly down the image, so interleaved
 * blocks give every GPU the same mix of short and long epipolar searches.  An incomplete last round is dealt from
 * context n-1 down whatever its parity (the rows next to the image border converge last: the context that holds the
 * first block of the image does not also get an extra block at the bottom).  Upload / download move the
 * owned rows only.  dmf_get_rows lists the owned image rows in local order (rows_out may be NULL to
 * query the count).
 */
int dmf_create_cyclic(const dmf_params *params, int device, int block_rows, int n_parts, int part, dmf_ctx **out);
int dmf_get_rows(const dmf_ctx *ctx, int *rows_out, int capacity, int *n_rows);
void dmf_destroy(dmf_ctx *ctx);

/* Geometry of the context. */
int dmf_get_params(const dmf_ctx *ctx, dmf_params *out);
int dmf_get_band(const dmf_ctx *ctx, int *row_begin, int *row_end);

/*
 * Set the reference image (`ref` argument of update(), :108; CV_8UC1, `step` bytes
 * per row as cv::Mat::step).  Also runs the once-per-reference precompute of the
 * reference-patch mean and centred energy (ref half of NCC(), :458-459,468,476).
 * _host: pageable or pinned host memory; _device: device memory on ctx's device.
 */
int dmf_set_reference(dmf_ctx *ctx, const uint8_t *ref_host, size_t step);
int dmf_set_reference_device(dmf_ctx *ctx, const uint8_t *ref_dev, size_t step);

/*
 * State maps `depth`, `depth_cov2` (CV_64F, :277-278).  Host pointers address the
 * FULL image (row 0), `step` in bytes; only the context's band rows are transferred.
 * dmf_fill_state is the device-side equivalent of Mat(h,w,CV_64F,init) (:277-278).
 * dmf_download_state synchronises the context stream first (update() results are
 * visible to the caller on return, :292-300).
 */
int dmf_fill_state(dmf_ctx *ctx, double init_depth, double init_cov2);
int dmf_upload_state(dmf_ctx *ctx, const double *depth, size_t depth_step,
                     const double *cov2, size_t cov2_step);
int dmf_download_state(dmf_ctx *ctx, double *depth, size_t depth_step,
                       double *cov2, size_t cov2_step);

/*
 * One update() call (:355-393) against the current frame `curr` (CV_8UC1) with
 * T_C_R given as Sophus stores it: unit quaternion (x,y,z,w) + translation.
 * Asynchronous on the context stream; the host frame is consumed (copied to a
 * device staging buffer) before return unless it is page-locked memory known to CUDA
 * (dmf_alloc_pinned(), cudaMallocHost, cudaHostRegister — detected with
 * cudaPointerGetAttributes), in which case it is copied straight from the caller's
 * buffer and must stay untouched until dmf_sync().
 * _device: frame already in HBM on ctx's device (e.g. received by NCCL broadcast);
 * the kernels that read it are ordered after everything previously enqueued on
 * `wait_stream` (a cudaStream_t).  With wait_stream == NULL the frame must be
 * complete in device memory when the call is made: the frame-only precompute runs
 * on an internal stream beside the previous update and is NOT ordered after work
 * enqueued on the context stream.  The frame must stay untouched until an event
 * recorded on the context stream after this call has completed (or dmf_sync()).
 */
int dmf_update(dmf_ctx *ctx, const uint8_t *curr_host, size_t step,
               const double q_xyzw[4], const double t_xyz[3]);
int dmf_update_device(dmf_ctx *ctx, const uint8_t *curr_dev, size_t step,
                      const double q_xyzw[4], const double t_xyz[3], void *wait_stream);

/*
 * STRICT drop-in form of update() (:355-393, call site :291) in one call: `depth` / `depth_cov2` (host, CV_64F) are
 * read, updated against `curr` and valid again on return, as the reference's caller reads them after every call
 * (:292-300).  The reference image is uploaded (and its patch statistics recomputed) only when its CONTENT changes
 * (64-bit hash); the maps are uploaded only when they differ from what the previous call returned (compared against a
 * pinned shadow copy); downloads go through pinned memory.  The context must own every interior row.
 */
int dmf_update_strict(dmf_ctx *ctx, const uint8_t *ref_host, size_t ref_step, const uint8_t *curr_host, size_t curr_step,
                      const double q_xyzw[4], const double t_xyz[3], double *depth, size_t depth_step,
                      double *depth_cov2, size_t cov2_step);

/*
 * The Gaussian fusion (:546-564) of an update is deferred: it runs inside the next
 * update's first kernel (its result feeds the next search straight from registers),
 * or as soon as anything reads or replaces the maps (every accessor below does so
 * implicitly).  dmf_flush() enqueues a pending fusion on the context stream without
 * waiting for it; dmf_sync() does the same and waits for all streams of the context.
 */
int dmf_flush(dmf_ctx *ctx);
int dmf_sync(dmf_ctx *ctx);

/*
 * Optional per-kernel timing for the roofline report (bench.py): with timing enabled every update()
 * serialises its kernels on the context stream and brackets them with CUDA events; the four slots are, in launch
 * order: set-up (advance_kernel = fusion of the previous update + set-up, or setup_kernel), block moments, ncc, and the
 * stand-alone fusion (only when one runs right after the update: debug planes on).
 * dmf_get_timing synchronises and returns the accumulated milliseconds per kernel class and the number
 * of frames they cover.  Off by default (the events cost ~1 % on small frames).
 */
int dmf_set_timing(dmf_ctx *ctx, int enable);
int dmf_get_timing(dmf_ctx *ctx, double ms_out[4], uint64_t *frames, int reset);

/* Work counters accumulated on the device; `reset` != 0 clears them after reading. Syncs. */
int dmf_read_counters(dmf_ctx *ctx, dmf_counters *out, int reset);

/*
 * Per-pixel decision flags of the LAST update (parity metric P2):
 * bit0 = passed the gate :366, bit1 = epipolarSearch accepted :443.  Disabled by
 * default; enabling costs one byte store per interior pixel per frame.
 */
int dmf_enable_flags(dmf_ctx *ctx, int enable);
int dmf_download_flags(dmf_ctx *ctx, uint8_t *flags_host, size_t step);
/* With flags enabled: per-pixel best NCC (ref:430-441) and (trip count of the loop ref:432) << 16 |
 * index of the winning iteration (0xFFFF: none) of the LAST update; dense W*H arrays. */
int dmf_download_debug(dmf_ctx *ctx, float *best_ncc_host, int32_t *samples_host);

/*
 * Raw device pointers for the multi-GPU gather (row-band sharding):
 * full-image pitch-linear maps, `pitch` bytes per row; rows outside the band hold the
 * last uploaded / filled values.  `stream` is the context's cudaStream_t.
 */
int dmf_device_state(dmf_ctx *ctx, double **depth_dev, double **cov2_dev, size_t *pitch);
int dmf_stream(dmf_ctx *ctx, void **stream);

/*
 * Device self-test: the kernels compute groups of IEEE-rounded FP64 quotients over one denominator with a shared
 * reciprocal (slamplay_b200/csrc/dmf_geometry.h quo2 / quo3).  Compares `n` pseudo-random groups (normal, huge, tiny,
 * zero, infinite, NaN and denormal operands) with __ddiv_rn bit for bit; *mismatches must come back 0.
 */
int dmf_selftest_division(int device, uint64_t n, uint64_t seed, uint64_t *mismatches);

/* Pinned host memory helpers for zero-staging dmf_update() calls.  dmf_host_register page-locks memory the caller
 * already owns (e.g. a POSIX shared-memory mapping of the frames that several ranks publish from, one frame each in
 * turn, so that the uploads use every GPU's PCIe link). */
int dmf_alloc_pinned(void **ptr, size_t bytes);
int dmf_free_pinned(void *ptr);
int dmf_host_register(void *ptr, size_t bytes);
int dmf_host_unregister(void *ptr);

/*
 * "Next" rows — consumers of the maps, evaluated on the device so a
 * caller need not download 16*W*H bytes per frame.
 *
 * dmf_evaluate_depth: evaludateDepth() (:569-590) over the context's band:
 *   sum of squared (truth - estimate) over interior pixels with variance < max_variance,
 *   and their count.  RMS = sqrt(sum_sq / count).  truth is a full-image host map.
 * dmf_variance_mask: getMaskFromVariance() (:199-204): mask = variance > max_variance ? 0 : 255
 *   (cv::threshold THRESH_BINARY_INV then convertTo CV_8U), band rows only.
 */
int dmf_set_truth(dmf_ctx *ctx, const double *truth_host, size_t step);
int dmf_evaluate_depth(dmf_ctx *ctx, double max_variance, double *sum_sq, uint64_t *count);
int dmf_variance_mask(dmf_ctx *ctx, double max_variance, uint8_t *mask_host, size_t step);

/*
 * getPointCloudFromImageAndDistance (utils/pointcloud/pointcloud_from_image_depth.h:42-89) as called at
 * ref:296-300: mask = getMaskFromVariance(depth_cov2, max_variance), distance = depth, T = identity.
 * `color_host`: the reference colour image (`channels` = 3 BGR as cv::imread gives it, or 1 / 4), full image.
 * Writes up to `capacity` points in the reference's scan order: xyz as 3 floats (PointXYZRGB narrows to
 * float), rgb as 3 bytes (r,g,b).  *n_points receives the number of valid points (may exceed capacity).
 * A context that owns part of the rows (band or block-cyclic) returns the points of its rows, in scan order.
 */
int dmf_point_cloud(dmf_ctx *ctx, const uint8_t *color_host, size_t color_step, int channels, double max_variance,
                    float *xyz_host, uint8_t *rgb_host, uint64_t capacity, uint64_t *n_points);

/*
 * Multi-GPU frame distribution on the copy engines (every rank needs the whole current frame).
 * One PRODUCER process (the rank that receives the frames) owns a ring of `n_slots` frames in its device memory;
 * every rank, the producer's included, opens the ring as CONSUMER number 0 .. n_consumers-1 from the 192-byte handle
 * (plain bytes: send them with any transport).  dmf_ring_publish enqueues the copy of one frame (pinned host memory or
 * device memory, `step` bytes per row) into the next slot; dmf_update_ring is dmf_update() against the next frame of
 * the ring: the context's copy stream waits for the slot, pulls it over NVLink peer-to-peer (or locally) with a copy
 * engine into the context's own buffer, releases the slot and launches the kernels.  All of it is stream-ordered
 * (cuStreamWaitValue32 / cuStreamWriteValue32 on a shared flag page): no host ever blocks on another, no SM is used
 * for the transfer, and frames are consumed in publication order by every consumer.  A consumer that never calls
 * dmf_update_ring stalls the producer after n_slots frames.
 */
#define DMF_RING_HANDLE_BYTES 192
typedef struct dmf_ring dmf_ring;
int dmf_ring_create(int device, int n_slots, int width, int height, int n_consumers, dmf_ring **out, uint8_t *handle_out);
int dmf_ring_open(int device, const uint8_t *handle, int consumer, dmf_ring **out);
void dmf_ring_close(dmf_ring *ring);
int dmf_ring_info(const dmf_ring *ring, int *n_slots, int *n_consumers, uint32_t *next_frame);
int dmf_ring_publish(dmf_ring *ring, const uint8_t *frame, size_t step, void *wait_stream);
int dmf_update_ring(dmf_ctx *ctx, dmf_ring *ring, const double q_xyzw[4], const double t_xyz[3]);

/*
 * Measured issue rates of the pipes the kernels of this path run on (the path is bound by instruction
 * issue and L1/TEX throughput; MEASURED_PEAKS.json holds HBM and bf16 only).  Micro-benchmarks, ~20 ms in total:
 * thread-operations per clock per SM (from clock64 of the slowest CTA), per second (CUDA events) and the SM clock the
 * two imply.  ldg64_l1 / ldg128_l1: coalesced loads that hit L1 — one warp-wide LDG.64 is 2 wavefronts of 128 bytes,
 * one LDG.128 is 4, so wavefronts/clk/SM = per_clk_sm / 32 * {2, 4}.  bench.py uses these as roofline denominators.
 */
typedef struct dmf_pipe_rate {
    double per_clk_sm;   /* thread-ops / clock / SM */
    double per_second;   /* thread-ops / s, whole chip */
    double eff_mhz;      /* SM clock during the measurement */
} dmf_pipe_rate;
typedef struct dmf_pipe_peaks_t {
    int32_t n_sm;
    int32_t reserved;
    dmf_pipe_rate ffma, dfma, idp4a, i2f_f64, ldg64_l1, ldg128_l1;
} dmf_pipe_peaks_t;
int dmf_pipe_peaks(int device, dmf_pipe_peaks_t *out);

#ifdef __cplusplus
}
#endif
#endif /* DMF_H_ */
