// ref_wrapper.cpp — C entry points onto the UNMODIFIED reference translation unit
// /root/reference/dense_mapping/test_monocular_mapping.cpp, compiled by oracle/Makefile
// against the stand-in headers of oracle/ref_shim/ into oracle/_ref/libdmf_ref.so.
// TEST INFRASTRUCTURE ONLY.  The reference hard-codes 640x480 and its intrinsics
// (ref:72-78), so these entry points exist for that geometry only.
#include <sophus/se3.hpp>
#include <Eigen/Core>
#include <opencv2/core/core.hpp>
#include <cstdint>

using Eigen::Vector2d;
using Sophus::SE3d;
using cv::Mat;

// Declarations of the reference's own functions (ref:107-112, 126-134, 146-152, 162).
void update(const Mat &ref, const Mat &curr, const SE3d &T_C_R, Mat &depth, Mat &depth_cov2);
bool epipolarSearch(const Mat &ref, const Mat &curr, const SE3d &T_C_R, const Vector2d &pt_ref, const double &depth_mu,
                    const double &depth_cov, Vector2d &pt_curr, Vector2d &epipolar_direction);
bool updateDepthFilter(const Vector2d &pt_ref, const Vector2d &pt_curr, const SE3d &T_C_R,
                       const Vector2d &epipolar_direction, Mat &depth, Mat &depth_cov2);
double NCC(const Mat &ref, const Mat &curr, const Vector2d &pt_ref, const Vector2d &pt_curr);

namespace {
const int W = 640, H = 480;
SE3d make(const double q[4], const double t[3]) {
    return SE3d::raw(Eigen::Quaterniond(q[3], q[0], q[1], q[2]), Eigen::Vector3d(t[0], t[1], t[2]));
}
}  // namespace

extern "C" {

int ref_width(void) { return W; }
int ref_height(void) { return H; }

void ref_update(const uint8_t *ref, size_t ref_step, const uint8_t *curr, size_t curr_step, const double q_xyzw[4],
                const double t_xyz[3], double *depth, size_t depth_step, double *cov2, size_t cov2_step) {
    Mat mref(H, W, cv::CV_8UC1, const_cast<uint8_t *>(ref), ref_step);
    Mat mcur(H, W, cv::CV_8UC1, const_cast<uint8_t *>(curr), curr_step);
    Mat md(H, W, cv::CV_64F, depth, depth_step), mc(H, W, cv::CV_64F, cov2, cov2_step);
    update(mref, mcur, make(q_xyzw, t_xyz), md, mc);
}

double ref_ncc(const uint8_t *ref, size_t ref_step, const uint8_t *curr, size_t curr_step, double rx, double ry,
               double cx, double cy) {
    Mat mref(H, W, cv::CV_8UC1, const_cast<uint8_t *>(ref), ref_step);
    Mat mcur(H, W, cv::CV_8UC1, const_cast<uint8_t *>(curr), curr_step);
    return NCC(mref, mcur, Vector2d(rx, ry), Vector2d(cx, cy));
}

// out = ok, pt_curr.x, pt_curr.y, dir.x, dir.y
int ref_epipolar_search(const uint8_t *ref, size_t ref_step, const uint8_t *curr, size_t curr_step, const double q[4],
                        const double t[3], double x, double y, double mu, double sigma, double out[5]) {
    Mat mref(H, W, cv::CV_8UC1, const_cast<uint8_t *>(ref), ref_step);
    Mat mcur(H, W, cv::CV_8UC1, const_cast<uint8_t *>(curr), curr_step);
    Vector2d pt(0, 0), dir(0, 0);
    bool ok = epipolarSearch(mref, mcur, make(q, t), Vector2d(x, y), mu, sigma, pt, dir);
    out[0] = ok; out[1] = pt[0]; out[2] = pt[1]; out[3] = dir[0]; out[4] = dir[1];
    return 0;
}

// Fuses one measurement into a 1-pixel-wide view of the maps; out = fused depth, fused cov2.
int ref_update_depth_filter(const double q[4], const double t[3], int rx, int ry, double cx, double cy, double dirx,
                            double diry, double depth_val, double cov2_val, double out[2]) {
    static thread_local double dbuf[W * H], cbuf[W * H];
    Mat md(H, W, cv::CV_64F, dbuf, W * sizeof(double)), mc(H, W, cv::CV_64F, cbuf, W * sizeof(double));
    dbuf[ry * W + rx] = depth_val;
    cbuf[ry * W + rx] = cov2_val;
    updateDepthFilter(Vector2d(rx, ry), Vector2d(cx, cy), make(q, t), Vector2d(dirx, diry), md, mc);
    out[0] = dbuf[ry * W + rx];
    out[1] = cbuf[ry * W + rx];
    return 0;
}

}  // extern "C"
