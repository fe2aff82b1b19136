// ref_wrapper.cpp — C entry points onto the UNMODIFIED reference translation unit
// /root/reference/dense_mapping/test_monocular_mapping.cpp, compiled by oracle/Makefile
// against the stand-in headers of oracle/ref_shim/ into oracle/_ref/libdmf_ref.so.
// TEST INFRASTRUCTURE ONLY.  The reference hard-codes 640x480 and its intrinsics
// (ref:72-78), so these entry points exist for that geometry only.
#include <sophus/se3.hpp>
#include <Eigen/Core>
#include <opencv2/core/core.hpp>
#include <pcl/point_types.h>
#include "camera/cam_utils.h"
#include "pointcloud/pointcloud_from_image_depth.h"  // the reference's own header (utils/pointcloud/)
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <sstream>
#include <string>
#include <vector>

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
// "next" rows: ref:569-590, ref:199-204, ref:317-352
void evaludateDepth(const Mat &depth_truth, const Mat &depth_estimate, const Mat &depth_variance, const double max_variance);
cv::Mat getMaskFromVariance(const cv::Mat &variance, const double max_variance);
bool readDatasetFiles(const std::string &path, std::vector<std::string> &color_image_files, std::vector<SE3d> &poses, cv::Mat &ref_depth);

namespace {
const int W = 640, H = 480;
SE3d make(const double q[4], const double t[3]) {
    return SE3d::raw(Eigen::Quaterniond(q[3], q[0], q[1], q[2]), Eigen::Vector3d(t[0], t[1], t[2]));
}
}  // namespace

extern "C" {

// 1 for the variant built from the translation unit with USE_INVERSE_DEPTH_FOR_FILTERING 1 (ref:63; oracle/Makefile)
int ref_inverse_depth(void) {
#ifdef REF_WRAPPER_INVERSE
    return 1;
#else
    return 0;
#endif
}

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

// evaludateDepth ref:569-590 prints "Average error (RMS) = x" on std::cout; the stream is captured at 17 significant
// digits so the printed number round-trips to the double the function computed.
double ref_evaluate_depth(const double *truth, size_t tstep, const double *est, size_t estep, const double *var, size_t vstep,
                          double max_variance) {
    Mat mt(H, W, cv::CV_64F, const_cast<double *>(truth), tstep), me(H, W, cv::CV_64F, const_cast<double *>(est), estep),
        mv(H, W, cv::CV_64F, const_cast<double *>(var), vstep);
    std::ostringstream ss;
    std::streambuf *old = std::cout.rdbuf(ss.rdbuf());
    const std::streamsize prec = std::cout.precision(17);
    evaludateDepth(mt, me, mv, max_variance);
    std::cout.precision(prec);
    std::cout.rdbuf(old);
    const std::string out = ss.str();
    const size_t eq = out.find('=');
    return eq == std::string::npos ? -1.0 : std::strtod(out.c_str() + eq + 1, nullptr);
}

// getMaskFromVariance ref:199-204 (cv::threshold THRESH_BINARY_INV + convertTo CV_8U from the OpenCV stand-in)
void ref_variance_mask(const double *var, size_t vstep, double max_variance, uint8_t *mask, size_t mstep) {
    Mat mv(H, W, cv::CV_64F, const_cast<double *>(var), vstep);
    Mat m = getMaskFromVariance(mv, max_variance);
    for (int y = 0; y < H; y++) std::memcpy(mask + size_t(y) * mstep, m.ptr<uint8_t>(y), W);
}

// getPointCloudFromImageAndDistance, utils/pointcloud/pointcloud_from_image_depth.h:42-89, called as at ref:296-300
// (BGR colour image, mask from getMaskFromVariance, border 20, T = identity).  Returns the number of points; writes at
// most `capacity` of them (x,y,z as the floats PointXYZRGB holds; r,g,b).
long long ref_point_cloud(const uint8_t *color_bgr, size_t color_step, const double *dist, size_t dstep, const uint8_t *mask,
                          size_t mstep, int border, float *xyz, uint8_t *rgb, long long capacity) {
    Mat mc(H, W, cv::CV_8UC3, const_cast<uint8_t *>(color_bgr), color_step);
    Mat md(H, W, cv::CV_64F, const_cast<double *>(dist), dstep);
    Mat mm(H, W, cv::CV_8UC1, const_cast<uint8_t *>(mask), mstep);
    const slamplay::Intrinsics K{481.2f, -480.0f, 319.5f, 239.5f};  // ref:75-78,282
    pcl::PointCloud<pcl::PointXYZRGB> cloud;
    slamplay::getPointCloudFromImageAndDistance<pcl::PointXYZRGB, double>(mc, md, mm, K, border, Eigen::Isometry3d::Identity(), cloud);
    long long n = 0;
    for (const auto &p : cloud.points) {
        if (n < capacity) {
            xyz[3 * n] = p.x; xyz[3 * n + 1] = p.y; xyz[3 * n + 2] = p.z;
            rgb[3 * n] = p.r; rgb[3 * n + 1] = p.g; rgb[3 * n + 2] = p.b;
        }
        n++;
    }
    return n;
}

// readDatasetFiles ref:317-352 on a REMODE-layout directory.  Returns the number of list entries (-1: failure);
// poses as the SE3d objects hold them after construction (unit quaternion x,y,z,w + translation), file names joined
// by '\n' into names (truncated to names_cap), the reference depth map (already / 100) into ref_depth (W*H doubles).
int ref_read_dataset(const char *path, double *poses7, int max_poses, char *names, size_t names_cap, double *ref_depth) {
    std::vector<std::string> files;
    std::vector<SE3d> poses;
    Mat depth;
    if (!readDatasetFiles(path, files, poses, depth)) return -1;
    std::string joined;
    for (size_t i = 0; i < files.size(); i++) { if (i) joined += '\n'; joined += files[i]; }
    if (names && names_cap) { std::strncpy(names, joined.c_str(), names_cap - 1); names[names_cap - 1] = 0; }
    for (size_t i = 0; i < poses.size() && (int)i < max_poses; i++) {
        const Eigen::Quaterniond &q = poses[i].unit_quaternion();
        const Eigen::Vector3d &t = poses[i].translation();
        double *o = poses7 + 7 * i;
        o[0] = q.x(); o[1] = q.y(); o[2] = q.z(); o[3] = q.w(); o[4] = t[0]; o[5] = t[1]; o[6] = t[2];
    }
    if (ref_depth) for (int y = 0; y < H; y++) std::memcpy(ref_depth + size_t(y) * W, depth.ptr<double>(y), W * sizeof(double));
    return (int)poses.size();
}

// T_C_R = T_WC(curr)^-1 * T_WC(ref) as composed at ref:289-290 from poses as readDatasetFiles builds them (ref:333-335)
void ref_compose_T_C_R(const double ref7[7], const double cur7[7], double out7[7]) {
    SE3d Tr(Eigen::Quaterniond(ref7[3], ref7[0], ref7[1], ref7[2]), Eigen::Vector3d(ref7[4], ref7[5], ref7[6]));
    SE3d Tc(Eigen::Quaterniond(cur7[3], cur7[0], cur7[1], cur7[2]), Eigen::Vector3d(cur7[4], cur7[5], cur7[6]));
    SE3d T = Tc.inverse() * Tr;
    const Eigen::Quaterniond &q = T.unit_quaternion();
    out7[0] = q.x(); out7[1] = q.y(); out7[2] = q.z(); out7[3] = q.w();
    out7[4] = T.translation()[0]; out7[5] = T.translation()[1]; out7[6] = T.translation()[2];
}

}  // extern "C"
