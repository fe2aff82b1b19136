// Inert stand-ins for PCL and the slamplay utils/ headers that only main() / the viz helpers
// of dense_mapping/test_monocular_mapping.cpp use.  TEST INFRASTRUCTURE ONLY; from scratch.
#pragma once
#include <cstdio>
#include <cstdlib>
#include <memory>
#include <vector>
#include "standin_eigen.h"
#include "standin_opencv.h"

namespace pcl {
struct PointXYZRGB { float x, y, z; unsigned char r, g, b; };
template <typename P> struct PointCloud {
    typedef std::shared_ptr<PointCloud<P>> Ptr;
    std::vector<P> points; unsigned width = 0, height = 0;
};
}  // namespace pcl

namespace slamplay {
struct Intrinsics { double fx, fy, cx, cy; };
template <typename PointT, typename ScalarT>
inline void getPointCloudFromImageAndDistance(const cv::Mat &, const cv::Mat &, const cv::Mat &, const Intrinsics &, int,
                                              const Eigen::Isometry3d &, pcl::PointCloud<PointT> &) {}
template <typename Cloud> struct PointCloudViz { void start() {} void update(const Cloud &) {} };
}  // namespace slamplay

#ifndef MSG_ASSERT
#define MSG_ASSERT(cond, msg) do { if (!(cond)) { std::fprintf(stderr, "assert failed: %s\n", #cond); std::abort(); } } while (0)
#endif
#ifndef STR
#define XSTR(x) #x
#define STR(x) XSTR(x)
#endif
