// Stand-ins for PCL's point-cloud container and for the slamplay viz helper that only main() of
// dense_mapping/test_monocular_mapping.cpp uses.  TEST INFRASTRUCTURE ONLY; from scratch.
// The reference's OWN headers on the path are NOT stood in for: utils/pointcloud/pointcloud_from_image_depth.h,
// utils/camera/cam_utils.h, utils/io/messages.h and utils/macros.h are compiled from /root/reference (oracle/Makefile
// puts $(REF_ROOT)/utils on the include path behind this directory).
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
    std::vector<P> points;
    unsigned width = 0, height = 0;
    void clear() { points.clear(); width = 0; height = 0; }
    size_t size() const { return points.size(); }
};
}  // namespace pcl

namespace slamplay {
template <typename Cloud> struct PointCloudViz { void start() {} void update(const Cloud &) {} };
}  // namespace slamplay
