// Stand-in for the subset of Eigen 3.3/3.4 used by the reference translation unit
// dense_mapping/test_monocular_mapping.cpp.  TEST INFRASTRUCTURE ONLY (see oracle/README.md).
// Written from scratch; it restates the ARITHMETIC ORDER of the Eigen operations the path
// uses (coefficient-wise expression evaluation, the unrolled reduction tree
// e0 + (e1 + e2) of a 3-vector dot product, guarded normalize(), Quaternion::_transformVector,
// ColPivHouseholderQR for a fixed 2x2 real matrix).  Eigen itself is not available in this image.
#pragma once
#include <cmath>
#include <limits>

namespace Eigen {

template <int N>
struct Vec {
    double v[N];
    Vec() {}
    Vec(double a, double b) { static_assert(N == 2, "2 coeffs"); v[0] = a; v[1] = b; }
    Vec(double a, double b, double c) { static_assert(N == 3, "3 coeffs"); v[0] = a; v[1] = b; v[2] = c; }
    double &operator()(int i, int = 0) { return v[i]; }
    const double &operator()(int i, int = 0) const { return v[i]; }
    double &operator[](int i) { return v[i]; }
    const double &operator[](int i) const { return v[i]; }
    double dot(const Vec &o) const {
        if (N == 2) return v[0] * o.v[0] + v[1] * o.v[1];
        return v[0] * o.v[0] + (v[1] * o.v[1] + v[2 % N] * o.v[2 % N]);  // redux_novec_unroller<0,3>: e0 + (e1 + e2)
    }
    double squaredNorm() const { return dot(*this); }
    double norm() const { return std::sqrt(squaredNorm()); }
    void normalize() {
        double z = squaredNorm();
        if (z > 0) { double n = std::sqrt(z); for (int i = 0; i < N; i++) v[i] /= n; }
    }
    Vec operator-() const { Vec r; for (int i = 0; i < N; i++) r.v[i] = -v[i]; return r; }
    Vec cross(const Vec &o) const {
        static_assert(N == 3, "cross needs 3 coeffs");
        return Vec(v[1] * o.v[2] - v[2] * o.v[1], v[2] * o.v[0] - v[0] * o.v[2], v[0] * o.v[1] - v[1] * o.v[0]);
    }
};
template <int N> inline Vec<N> operator+(const Vec<N> &a, const Vec<N> &b) { Vec<N> r; for (int i = 0; i < N; i++) r.v[i] = a.v[i] + b.v[i]; return r; }
template <int N> inline Vec<N> operator-(const Vec<N> &a, const Vec<N> &b) { Vec<N> r; for (int i = 0; i < N; i++) r.v[i] = a.v[i] - b.v[i]; return r; }
template <int N> inline Vec<N> operator*(const Vec<N> &a, double s) { Vec<N> r; for (int i = 0; i < N; i++) r.v[i] = a.v[i] * s; return r; }
template <int N> inline Vec<N> operator*(double s, const Vec<N> &a) { Vec<N> r; for (int i = 0; i < N; i++) r.v[i] = s * a.v[i]; return r; }
template <int N> inline Vec<N> operator/(const Vec<N> &a, double s) { Vec<N> r; for (int i = 0; i < N; i++) r.v[i] = a.v[i] / s; return r; }

typedef Vec<2> Vector2d;
typedef Vec<3> Vector3d;

struct Matrix2d {
    double m[2][2];
    double &operator()(int r, int c) { return m[r][c]; }
    const double &operator()(int r, int c) const { return m[r][c]; }
};

struct Quaterniond {
    double x_, y_, z_, w_;
    Quaterniond() : x_(0), y_(0), z_(0), w_(1) {}
    Quaterniond(double w, double x, double y, double z) : x_(x), y_(y), z_(z), w_(w) {}
    double x() const { return x_; } double y() const { return y_; } double z() const { return z_; } double w() const { return w_; }
    double norm() const { return std::sqrt((x_ * x_ + z_ * z_) + (y_ * y_ + w_ * w_)); }  // packet reduction of (x,y,z,w)
    void normalize() { double n = norm(); x_ /= n; y_ /= n; z_ /= n; w_ /= n; }
    Quaterniond conjugate() const { return Quaterniond(w_, -x_, -y_, -z_); }
    Vector3d _transformVector(const Vector3d &v) const {
        Vector3d q(x_, y_, z_);
        Vector3d uv = q.cross(v);
        uv = uv + uv;
        return v + w_ * uv + q.cross(uv);
    }
};

// Transform<double,3,Isometry>: 3x3 linear part + translation; T * p = linear * p + translation
struct Isometry3d {
    double m[3][3], t[3];
    static Isometry3d Identity() {
        Isometry3d r;
        for (int i = 0; i < 3; i++) { for (int j = 0; j < 3; j++) r.m[i][j] = (i == j) ? 1.0 : 0.0; r.t[i] = 0.0; }
        return r;
    }
    Vector3d operator*(const Vector3d &p) const {
        return Vector3d((m[0][0] * p[0] + m[0][1] * p[1]) + m[0][2] * p[2] + t[0], (m[1][0] * p[0] + m[1][1] * p[1]) + m[1][2] * p[2] + t[1],
                        (m[2][0] * p[0] + m[2][1] * p[1]) + m[2][2] * p[2] + t[2]);
    }
};
template <int N> inline Vec<N> &operator*=(Vec<N> &a, double s) { for (int i = 0; i < N; i++) a.v[i] *= s; return a; }

// ColPivHouseholderQR<Matrix2d>: computeInPlace() + _solve_impl() of Eigen 3.3/3.4, fixed 2x2 real.
template <typename M> class ColPivHouseholderQR;
template <>
class ColPivHouseholderQR<Matrix2d> {
public:
    explicit ColPivHouseholderQR(const Matrix2d &A) { qr_ = A; compute(); }
    Vector2d solve(const Vector2d &b) const {
        Vector2d x(0, 0);
        if (nonzero_pivots_ == 0) return x;
        double c[2] = {b[0], b[1]};
        if (hcoeff_[0] != 0) {
            double ess = qr_.m[1][0];
            double tmp = ess * c[1];
            tmp += c[0];
            c[0] -= hcoeff_[0] * tmp;
            c[1] -= hcoeff_[0] * ess * tmp;
        }
        if (nonzero_pivots_ == 2) {
            c[1] = c[1] / qr_.m[1][1];
            c[0] = (c[0] - qr_.m[0][1] * c[1]) / qr_.m[0][0];
        } else {
            c[0] = c[0] / qr_.m[0][0];
        }
        for (int i = 0; i < nonzero_pivots_; i++) x[perm_[i]] = c[i];
        return x;
    }

private:
    void compute() {
        const double eps = std::numeric_limits<double>::epsilon();
        double (*m)[2] = qr_.m;
        double upd[2], dir[2];
        for (int k = 0; k < 2; k++) { dir[k] = std::sqrt(m[0][k] * m[0][k] + m[1][k] * m[1][k]); upd[k] = dir[k]; }
        const double maxn = upd[0] >= upd[1] ? upd[0] : upd[1];
        const double threshold_helper = (maxn * eps) * (maxn * eps) / 2.0;
        const double downdate = std::sqrt(eps);
        nonzero_pivots_ = 2;
        int transp[2] = {0, 1};
        hcoeff_[0] = hcoeff_[1] = 0;
        for (int k = 0; k < 2; k++) {
            int big = k;
            for (int j = k + 1; j < 2; j++) if (upd[j] > upd[big]) big = j;
            double big_sq = upd[big] * upd[big];
            if (nonzero_pivots_ == 2 && big_sq < threshold_helper * double(2 - k)) nonzero_pivots_ = k;
            transp[k] = big;
            if (k != big) {
                for (int r = 0; r < 2; r++) { double t = m[r][k]; m[r][k] = m[r][big]; m[r][big] = t; }
                double t = upd[k]; upd[k] = upd[big]; upd[big] = t;
                t = dir[k]; dir[k] = dir[big]; dir[big] = t;
            }
            double c0 = m[k][k];
            double tail_sq = (k == 0) ? m[1][0] * m[1][0] : 0.0;
            double tau, beta;
            if (tail_sq <= std::numeric_limits<double>::min()) { tau = 0; beta = c0; if (k == 0) m[1][0] = 0; }
            else {
                beta = std::sqrt(c0 * c0 + tail_sq);
                if (c0 >= 0) beta = -beta;
                m[1][0] = m[1][0] / (c0 - beta);
                tau = (beta - c0) / beta;
            }
            m[k][k] = beta;
            hcoeff_[k] = tau;
            if (k == 0) {
                if (tau != 0) {
                    double ess = m[1][0];
                    double tmp = ess * m[1][1];
                    tmp += m[0][1];
                    m[0][1] -= tau * tmp;
                    m[1][1] -= tau * ess * tmp;
                }
                if (upd[1] != 0) {
                    double temp = std::fabs(m[0][1]) / upd[1];
                    temp = (1.0 + temp) * (1.0 - temp);
                    temp = temp < 0 ? 0 : temp;
                    double r = upd[1] / dir[1];
                    double temp2 = temp * (r * r);
                    if (temp2 <= downdate) { dir[1] = std::fabs(m[1][1]); upd[1] = dir[1]; }
                    else upd[1] *= std::sqrt(temp);
                }
            }
        }
        perm_[0] = 0; perm_[1] = 1;
        for (int k = 0; k < 2; k++) { int t = perm_[k]; perm_[k] = perm_[transp[k]]; perm_[transp[k]] = t; }
    }
    Matrix2d qr_;
    double hcoeff_[2];
    int perm_[2];
    int nonzero_pivots_;
};

}  // namespace Eigen
