// Stand-in for the subset of Sophus (strasdat/Sophus @61f9a98, build_thirdparty.sh:163-167)
// used by dense_mapping/test_monocular_mapping.cpp: SE3d as unit quaternion + translation.
// TEST INFRASTRUCTURE ONLY; written from scratch (Sophus is not available in this image).
#pragma once
#include "../standin_eigen.h"

namespace Sophus {

class SO3d {
public:
    SO3d() {}
    explicit SO3d(const Eigen::Quaterniond &q) : q_(q) { q_.normalize(); }  // SO3 ctor normalises
    static SO3d raw(const Eigen::Quaterniond &q) { SO3d r; r.q_ = q; return r; }  // test hook: no normalisation
    const Eigen::Quaterniond &unit_quaternion() const { return q_; }
    SO3d inverse() const { return SO3d(q_.conjugate()); }
    Eigen::Vector3d operator*(const Eigen::Vector3d &p) const { return q_._transformVector(p); }
    SO3d operator*(const SO3d &o) const {
        const Eigen::Quaterniond &a = q_, &b = o.q_;
        return SO3d(Eigen::Quaterniond(a.w() * b.w() - a.x() * b.x() - a.y() * b.y() - a.z() * b.z(),
                                       a.w() * b.x() + a.x() * b.w() + a.y() * b.z() - a.z() * b.y(),
                                       a.w() * b.y() + a.y() * b.w() + a.z() * b.x() - a.x() * b.z(),
                                       a.w() * b.z() + a.z() * b.w() + a.x() * b.y() - a.y() * b.x()));
    }
private:
    Eigen::Quaterniond q_;
};

class SE3d {
public:
    SE3d() : t_(0, 0, 0) {}
    SE3d(const Eigen::Quaterniond &q, const Eigen::Vector3d &t) : so3_(q), t_(t) {}
    SE3d(const SO3d &r, const Eigen::Vector3d &t) : so3_(r), t_(t) {}
    static SE3d raw(const Eigen::Quaterniond &q, const Eigen::Vector3d &t) { return SE3d(SO3d::raw(q), t); }
    const SO3d &so3() const { return so3_; }
    const Eigen::Vector3d &translation() const { return t_; }
    const Eigen::Quaterniond &unit_quaternion() const { return so3_.unit_quaternion(); }
    SE3d inverse() const { SO3d invR = so3_.inverse(); return SE3d(invR, invR * (t_ * -1.0)); }
    Eigen::Vector3d operator*(const Eigen::Vector3d &p) const { return so3_ * p + t_; }
    SE3d operator*(const SE3d &o) const { return SE3d(so3_ * o.so3_, t_ + so3_ * o.t_); }
private:
    SO3d so3_;
    Eigen::Vector3d t_;
};

}  // namespace Sophus
