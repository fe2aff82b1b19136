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

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <map>
#include <memory>
#include <string>
#include <tuple>
#include <vector>

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
// SE3d(Quaterniond(qw,qx,qy,qz), Vector3d(tx,ty,tz)) ref:333-335
inline SE3d make_pose(double qx, double qy, double qz, double qw, double tx, double ty, double tz) {
    return SE3d(Eigen::Quaterniond(qw, qx, qy, qz), Eigen::Vector3d(tx, ty, tz));
}
// pose_curr_TWC.inverse() * pose_ref_TWC ref:289-290
inline SE3d relative_pose(const SE3d &T_WC_ref, const SE3d &T_WC_curr) { return T_WC_curr.inverse() * T_WC_ref; }
inline Mat make_mat64(int rows, int cols) { return Mat(rows, cols, CV_64F); }
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
    std::shared_ptr<std::vector<unsigned char>> own_;  // set when the Mat owns its pixels (make_mat64, like cv::Mat(rows, cols, type))
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
inline Mat make_mat64(int rows, int cols) {
    Mat m;
    m.own_ = std::make_shared<std::vector<unsigned char>>(size_t(rows) * cols * sizeof(double));
    m.data = m.own_->data(); m.step = size_t(cols) * sizeof(double); m.rows = rows; m.cols = cols; m.type_ = kType64F;
    return m;
}
// The Sophus / Eigen arithmetic the reference's driver applies to poses, in the same operation order (so a build
// without Sophus hands the kernels the same bits): quaternion normalisation of the SE3d constructor
// (packet reduction (x2+z2)+(y2+w2)), Quaternion::_transformVector, SE3::inverse, SE3 * SE3.
namespace detail {
inline void qnormalize(double q[4]) {
    const double n = std::sqrt((q[0] * q[0] + q[2] * q[2]) + (q[1] * q[1] + q[3] * q[3]));
    for (int i = 0; i < 4; i++) q[i] /= n;
}
inline void cross(const double a[3], const double b[3], double o[3]) {
    o[0] = a[1] * b[2] - a[2] * b[1]; o[1] = a[2] * b[0] - a[0] * b[2]; o[2] = a[0] * b[1] - a[1] * b[0];
}
inline void rotate(const double q[4], const double v[3], double o[3]) {
    double uv[3], c[3];
    cross(q, v, uv);
    for (int i = 0; i < 3; i++) uv[i] = uv[i] + uv[i];
    cross(q, uv, c);
    for (int i = 0; i < 3; i++) o[i] = v[i] + uv[i] * q[3] + c[i];
}
}  // namespace detail
// SE3d(Quaterniond(qw,qx,qy,qz), Vector3d(tx,ty,tz)) ref:333-335 — the constructor normalises the quaternion
inline SE3d make_pose(double qx, double qy, double qz, double qw, double tx, double ty, double tz) {
    SE3d T;
    T.q[0] = qx; T.q[1] = qy; T.q[2] = qz; T.q[3] = qw;
    detail::qnormalize(T.q);
    T.t[0] = tx; T.t[1] = ty; T.t[2] = tz;
    return T;
}
// pose_curr_TWC.inverse() * pose_ref_TWC ref:289-290
inline SE3d relative_pose(const SE3d &T_WC_ref, const SE3d &T_WC_curr) {
    SE3d inv;  // Sophus SE3::inverse: invR = conj(q) (normalised), t' = invR * (t * -1)
    inv.q[0] = -T_WC_curr.q[0]; inv.q[1] = -T_WC_curr.q[1]; inv.q[2] = -T_WC_curr.q[2]; inv.q[3] = T_WC_curr.q[3];
    detail::qnormalize(inv.q);
    const double nt[3] = {T_WC_curr.t[0] * -1.0, T_WC_curr.t[1] * -1.0, T_WC_curr.t[2] * -1.0};
    detail::rotate(inv.q, nt, inv.t);
    SE3d out;  // SE3 * SE3: (R_a R_b (normalised), t_a + R_a t_b)
    const double *a = inv.q, *b = T_WC_ref.q;
    out.q[0] = a[3] * b[0] + a[0] * b[3] + a[1] * b[2] - a[2] * b[1];
    out.q[1] = a[3] * b[1] + a[1] * b[3] + a[2] * b[0] - a[0] * b[2];
    out.q[2] = a[3] * b[2] + a[2] * b[3] + a[0] * b[1] - a[1] * b[0];
    out.q[3] = a[3] * b[3] - a[0] * b[0] - a[1] * b[1] - a[2] * b[2];
    detail::qnormalize(out.q);
    double r[3];
    detail::rotate(inv.q, T_WC_ref.t, r);
    for (int i = 0; i < 3; i++) out.t[i] = inv.t[i] + r[i];
    return out;
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

// readDatasetFiles (ref:317-352): the REMODE test-set layout.
//   <path>/first_200_frames_traj_over_table_input_sequence.txt   one line per frame: image tx ty tz qx qy qz qw (T_WC)
//   <path>/images/<image>                                        (ref:332)
//   <path>/depthmaps/scene_000.depth                             width*height numbers in centimetres (ref:341-349: / 100)
// Same signature and results as the reference's function, with one deliberate difference: the reference's
// `while (!fin.eof())` loop appends one bogus entry (empty file name, uninitialised pose) when the list ends with a
// newline, which its driver then skips because imread fails (ref:288); this reader returns the complete entries only.
inline bool readDatasetFiles(const std::string &path, std::vector<std::string> &color_image_files, std::vector<SE3d> &poses,
                             Mat &ref_depth, int width = 640, int height = 480) {
    std::ifstream fin(path + "/first_200_frames_traj_over_table_input_sequence.txt");
    if (!fin) return false;
    for (;;) {
        std::string image;
        double d[7];
        if (!(fin >> image)) break;
        bool ok = true;
        for (double &v : d) ok = ok && bool(fin >> v);
        if (!ok) break;
        color_image_files.push_back(path + std::string("/images/") + image);
        poses.push_back(make_pose(d[3], d[4], d[5], d[6], d[0], d[1], d[2]));
    }
    fin.close();
    fin.open(path + "/depthmaps/scene_000.depth");
    ref_depth = make_mat64(height, width);
    if (!fin) return false;
    for (int y = 0; y < height; y++)
        for (int x = 0; x < width; x++) {
            double depth = 0;
            fin >> depth;
            ref_depth.template ptr<double>(y)[x] = depth / 100.0;
        }
    return true;
}

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
    // One strict update() (:355): maps read from and valid again in the caller's memory; the reference image is
    // uploaded only when its content changes, unchanged maps are not uploaded again (dmf_update_strict).
    void updateStrict(const Mat &ref, const Mat &curr, const SE3d &T_C_R, Mat &depth, Mat &depth_cov2) {
        check8(ref, "ref"); check8(curr, "curr"); check64(depth, "depth"); check64(depth_cov2, "depth_cov2");
        double q[4], t[3];
        pose_of(T_C_R, q, t);
        if (dmf_update_strict(ctx_, ref.data, ref.step, curr.data, curr.step, q, t, reinterpret_cast<double *>(depth.data), depth.step,
                              reinterpret_cast<double *>(depth_cov2.data), depth_cov2.step) != DMF_OK) die("dmf_update_strict", ctx_);
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
    it->second->updateStrict(ref, curr, T_C_R, depth, depth_cov2);
}

}  // namespace slamplay_b200
