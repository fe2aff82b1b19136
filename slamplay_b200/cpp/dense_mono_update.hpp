// dense_mono_update.hpp — drop-in C++ host shim for the reference's call surface.
//
//   void update(const Mat &ref, const Mat &curr, const SE3d &T_C_R, Mat &depth, Mat &depth_cov2);
//   (luigifreda/slamplay dense_mapping/test_monocular_mapping.cpp:107-112, :355-393; call site :291)
//
// Header-only.  Works on cv::Mat / Sophus::SE3d when those headers are available (define
// DMF_USE_OPENCV_SOPHUS or let __has_include find them), otherwise on the layout-compatible stand-ins
// below, which expose exactly the members the path uses: Mat::{data, step, rows, cols, type()} and
// SE3d::{unit_quaternion(), translation()}.  All computation is forwarded to the C ABI of
// include/dmf.h (libdmf.so, CUDA sm_100a).  There is no CPU fallback: errors abort with a message,
// mirroring MSG_ASSERT (utils/io/messages.h:71-83) since update() returns void.
//
// Two modes:
//   slamplay_b200::update(ref, curr, T_C_R, depth, depth_cov2)     STRICT: maps uploaded, updated and
//        downloaded inside the call — the reference's caller reads them after every call (:292-300).
//   slamplay_b200::DenseMonoMapper                                   RESIDENT: maps stay in HBM for the
//        whole sequence; only the u8 frame and the pose cross PCIe per update; download() on demand.
#pragma once

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <tuple>

#include "../../include/dmf.h"

#if defined(DMF_USE_OPENCV_SOPHUS) || (__has_include(<opencv2/core/core.hpp>) && __has_include(<sophus/se3.hpp>))
#include <opencv2/core/core.hpp>
#include <sophus/se3.hpp>
namespace slamplay_b200 {
using Mat = cv::Mat;
using SE3d = Sophus::SE3d;
constexpr int kType8UC1 = CV_8UC1;
constexpr int kType64F = CV_64F;
inline void pose_of(const SE3d &T, double q[4], double t[3]) {
    const auto &uq = T.unit_quaternion();
    q[0] = uq.x(); q[1] = uq.y(); q[2] = uq.z(); q[3] = uq.w();
    const auto &tr = T.translation();
    t[0] = tr[0]; t[1] = tr[1]; t[2] = tr[2];
}
}  // namespace slamplay_b200
#else
namespace slamplay_b200 {
constexpr int kType8UC1 = 0;  // CV_8UC1
constexpr int kType64F = 6;   // CV_64F
// Stand-in for the cv::Mat members the path touches (non-owning view).
struct Mat {
    unsigned char *data = nullptr;
    size_t step = 0;
    int rows = 0, cols = 0;
    int type_ = kType8UC1;
    Mat() = default;
    Mat(int r, int c, int type, void *ext, size_t ext_step) : data(static_cast<unsigned char *>(ext)), step(ext_step), rows(r), cols(c), type_(type) {}
    int type() const { return type_; }
    template <typename T> T *ptr(int r) { return reinterpret_cast<T *>(data + size_t(r) * step); }
    template <typename T> const T *ptr(int r) const { return reinterpret_cast<const T *>(data + size_t(r) * step); }
};
// Stand-in for Sophus::SE3d as the path sees it: unit quaternion (x,y,z,w) + translation.
struct SE3d {
    double q[4] = {0, 0, 0, 1};
    double t[3] = {0, 0, 0};
};
inline void pose_of(const SE3d &T, double q[4], double t[3]) {
    std::memcpy(q, T.q, sizeof(T.q));
    std::memcpy(t, T.t, sizeof(T.t));
}
}  // namespace slamplay_b200
#endif

namespace slamplay_b200 {

[[noreturn]] inline void die(const char *what, const dmf_ctx *ctx) {
    std::fprintf(stderr, "slamplay_b200: %s: %s\n", what, dmf_last_error(ctx));
    std::abort();
}

#define DMF_ASSERT(cond, msg)                                                       \
    do {                                                                            \
        if (!(cond)) {                                                              \
            std::fprintf(stderr, "slamplay_b200: assertion failed: %s (%s)\n", msg, #cond); \
            std::abort();                                                           \
        }                                                                           \
    } while (0)

// Resident mapper: one context per image geometry / device.
class DenseMonoMapper {
public:
    explicit DenseMonoMapper(const dmf_params &p, int device = 0) : p_(p) {
        if (dmf_create(&p_, device, 0, p_.height, &ctx_) != DMF_OK) die("dmf_create", nullptr);
    }
    DenseMonoMapper(int width, int height, int device = 0) {
        if (dmf_default_params(&p_, width, height, 0) != DMF_OK) die("dmf_default_params", nullptr);
        if (dmf_create(&p_, device, 0, p_.height, &ctx_) != DMF_OK) die("dmf_create", nullptr);
    }
    ~DenseMonoMapper() { dmf_destroy(ctx_); }
    DenseMonoMapper(const DenseMonoMapper &) = delete;
    DenseMonoMapper &operator=(const DenseMonoMapper &) = delete;

    const dmf_params &params() const { return p_; }

    void setReference(const Mat &ref) {
        check8(ref, "ref");
        if (dmf_set_reference(ctx_, ref.data, ref.step) != DMF_OK) die("dmf_set_reference", ctx_);
    }
    // Mat depth(height, width, CV_64F, init_depth), depth_cov2(..., init_cov2)  (:270-278)
    void init(double init_depth = 3.0, double init_cov2 = 3.0) {
        DMF_ASSERT(init_cov2 < p_.max_cov, "Please increase max_cov above the init cov");  // :276
        if (dmf_fill_state(ctx_, init_depth, init_cov2) != DMF_OK) die("dmf_fill_state", ctx_);
    }
    void upload(const Mat &depth, const Mat &depth_cov2) {
        check64(depth, "depth"); check64(depth_cov2, "depth_cov2");
        if (dmf_upload_state(ctx_, reinterpret_cast<const double *>(depth.data), depth.step,
                             reinterpret_cast<const double *>(depth_cov2.data), depth_cov2.step) != DMF_OK) die("dmf_upload_state", ctx_);
    }
    void update(const Mat &curr, const SE3d &T_C_R) {
        check8(curr, "curr");
        double q[4], t[3];
        pose_of(T_C_R, q, t);
        if (dmf_update(ctx_, curr.data, curr.step, q, t) != DMF_OK) die("dmf_update", ctx_);
    }
    void download(Mat &depth, Mat &depth_cov2) {
        check64(depth, "depth"); check64(depth_cov2, "depth_cov2");
        if (dmf_download_state(ctx_, reinterpret_cast<double *>(depth.data), depth.step,
                               reinterpret_cast<double *>(depth_cov2.data), depth_cov2.step) != DMF_OK) die("dmf_download_state", ctx_);
    }
    dmf_counters counters(bool reset = false) {
        dmf_counters c{};
        if (dmf_read_counters(ctx_, &c, reset ? 1 : 0) != DMF_OK) die("dmf_read_counters", ctx_);
        return c;
    }
    dmf_ctx *ctx() { return ctx_; }

private:
    void check8(const Mat &m, const char *name) const {
        DMF_ASSERT(m.data != nullptr && m.type() == kType8UC1, name);
        DMF_ASSERT(m.cols == p_.width && m.rows == p_.height, "Shoud be equal to the one set above!");  // :265
    }
    void check64(const Mat &m, const char *name) const {
        DMF_ASSERT(m.data != nullptr && m.type() == kType64F, name);
        DMF_ASSERT(m.cols == p_.width && m.rows == p_.height, name);
    }
    dmf_params p_{};
    dmf_ctx *ctx_ = nullptr;
};

// STRICT drop-in with the reference's exact signature (:107-112).  A mapper per image size is cached.
inline void update(const Mat &ref, const Mat &curr, const SE3d &T_C_R, Mat &depth, Mat &depth_cov2) {
    static std::map<std::tuple<int, int>, std::unique_ptr<DenseMonoMapper>> cache;
    auto key = std::make_tuple(ref.cols, ref.rows);
    auto it = cache.find(key);
    if (it == cache.end()) it = cache.emplace(key, std::make_unique<DenseMonoMapper>(ref.cols, ref.rows)).first;
    DenseMonoMapper &m = *it->second;
    m.setReference(ref);
    m.upload(depth, depth_cov2);
    m.update(curr, T_C_R);
    m.download(depth, depth_cov2);
}

}  // namespace slamplay_b200
