// =============================================================================
// dense_mono_oracle.cpp — CPU ORACLE.  TEST INFRASTRUCTURE ONLY.
//
// A from-scratch FP64 restatement of the dense monocular depth-filter hot path of
// luigifreda/slamplay, dense_mapping/test_monocular_mapping.cpp (cited below as
// "ref:LINE").  It exists to CHECK the CUDA path (tests/, __graft_entry__.smoke(),
// bench.py's cpu_baseline / --impl reference legs).  Nothing under slamplay_b200/
// may link, import or execute it: the product path has no CPU fallback.
//
// Pinning status.  The reference holds no golden vectors / tests for this path
//  and its translation unit needs Eigen, Sophus, OpenCV, PCL and
// Pangolin, none of which exist in this image.  The pin we do have: oracle/Makefile
// compiles the UNMODIFIED reference translation unit from /root/reference against
// the minimal stand-in headers in oracle/ref_shim/ (which restate only the
// third-party types, not the path) into oracle/_ref/, and tests/test_oracle_vs_ref.py
// checks this restatement against it bit for bit at the reference's fixed 640x480
// geometry.  Frozen outputs of that build live in tests/golden/.
//
// What is followed, in the reference's operation order:
//   constants                        ref:72-89   (runtime dmf_params here)
//   getBilinearInterpolatedValue     ref:165-174
//   px2cam / cam2px / inside         ref:207-224
//   update                           ref:355-393
//   epipolarSearch                   ref:397-447 (accumulated l += 0.7, first strict max)
//   NCC                              ref:449-480 (two-pass, dy-outer / dx-inner, FP64)
//   updateDepthFilter                ref:482-567 (both USE_INVERSE_DEPTH_FOR_FILTERING arms)
//   evaludateDepth                   ref:569-590
//   getMaskFromVariance              ref:199-204
//   pose chain T_C_R                 ref:289-290,333-335
//   getPointCloudFromImageAndDistance  utils/pointcloud/pointcloud_from_image_depth.h:42-89
// Third-party arithmetic restated from the upstream projects (not vendored in the
// reference): Sophus @61f9a98 SE3d (unit quaternion + translation; rotate via
// Eigen's Quaternion::_transformVector), Eigen 3.3/3.4 normalize() (guarded
// against a zero norm) and ColPivHouseholderQR<Matrix2d>::solve.
//
// Build: see oracle/Makefile (flags of the reference's release build,
// CMakeLists.txt:52-54,66: -O3 -march=native -fopenmp; -ffp-contract=off is added
// so that results do not depend on the host's FMA support).
// =============================================================================
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <vector>

#include "../include/dmf.h"

namespace {

struct V2 { double x, y; };
struct V3 { double x, y, z; };

inline V3 operator*(const V3 &a, double s) { return {a.x * s, a.y * s, a.z * s}; }
inline V3 operator+(const V3 &a, const V3 &b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
// Eigen's unrolled reduction of a 3-vector: e0 + (e1 + e2)  (redux_novec_unroller<0,3>)
inline double dot(const V3 &a, const V3 &b) { return a.x * b.x + (a.y * b.y + a.z * b.z); }
inline V3 cross(const V3 &a, const V3 &b) {
    return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
inline double norm(const V3 &a) { return std::sqrt(dot(a, a)); }
// Eigen >= 3.3 MatrixBase::normalize(): divides only when squaredNorm() > 0.
inline void normalize(V3 &a) {
    double z = dot(a, a);
    if (z > 0) { double n = std::sqrt(z); a.x /= n; a.y /= n; a.z /= n; }
}
inline void normalize(V2 &a) {
    double z = a.x * a.x + a.y * a.y;
    if (z > 0) { double n = std::sqrt(z); a.x /= n; a.y /= n; }
}

// Sophus::SE3d stand-in: unit quaternion (x,y,z,w) + translation.
struct Quat { double x, y, z, w; };
struct SE3 { Quat q; V3 t; };

// Eigen QuaternionBase::_transformVector: v + w*(2 q×v) + q×(2 q×v)
inline V3 rotate(const Quat &q, const V3 &v) {
    V3 qv{q.x, q.y, q.z};
    V3 uv = cross(qv, v);
    uv = uv + uv;
    return v + uv * q.w + cross(qv, uv);
}
inline V3 apply(const SE3 &T, const V3 &p) { return rotate(T.q, p) + T.t; }  // Sophus SE3 * point

inline Quat normalized(const Quat &q) {  // Sophus SO3(Quaternion) constructor: coeffs /= norm
    double n = std::sqrt((q.x * q.x + q.z * q.z) + (q.y * q.y + q.w * q.w));  // packet reduction of (x,y,z,w)
    return {q.x / n, q.y / n, q.z / n, q.w / n};
}
inline SE3 inverse(const SE3 &T) {  // Sophus SE3::inverse: invR = so3().inverse(); (invR, invR * (t * -1))
    Quat c = normalized(Quat{-T.q.x, -T.q.y, -T.q.z, T.q.w});
    V3 nt{T.t.x * -1.0, T.t.y * -1.0, T.t.z * -1.0};
    return {c, rotate(c, nt)};
}
inline Quat qmul(const Quat &a, const Quat &b) {  // Sophus SO3 * SO3 (explicit Hamilton product), then normalised by the ctor
    Quat r{a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y,
           a.w * b.y + a.y * b.w + a.z * b.x - a.x * b.z,
           a.w * b.z + a.z * b.w + a.x * b.y - a.y * b.x,
           a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z};
    return normalized(r);
}
inline SE3 compose(const SE3 &A, const SE3 &B) {  // Sophus SE3 * SE3: (R_a R_b, t_a + R_a t_b)
    return {qmul(A.q, B.q), A.t + rotate(A.q, B.t)};
}

struct Cam {
    int width, height, border;
    double fx, fy, cx, cy;
    double step, max_half_len, min_depth, n_sigma, ncc_thresh, min_cov, max_cov;
    int inverse_depth;
};

inline Cam make_cam(const dmf_params &p) {
    return {p.width, p.height, p.border, p.fx, p.fy, p.cx, p.cy, p.step, p.max_half_len,
            p.min_depth, p.n_sigma, p.ncc_thresh, p.min_cov, p.max_cov, p.inverse_depth};
}

// ref:207-212
inline V3 px2cam(const Cam &c, const V2 &px) { return {(px.x - c.cx) / c.fx, (px.y - c.cy) / c.fy, 1.0}; }
// ref:215-219 (no Z > 0 check)
inline V2 cam2px(const Cam &c, const V3 &p) { return {p.x * c.fx / p.z + c.cx, p.y * c.fy / p.z + c.cy}; }
// ref:222-224 — asymmetric on purpose: '<' for x, '<=' for y
inline bool inside(const Cam &c, const V2 &pt) {
    return pt.x >= c.border && pt.y >= c.border && pt.x + c.border < c.width && pt.y + c.border <= c.height;
}

struct Img8 { const uint8_t *data; size_t step; };

// ref:165-174
inline double bilinear(const Img8 &img, double px, double py) {
    const uint8_t *d = &img.data[size_t(int(py)) * img.step + size_t(int(px))];
    double xx = px - std::floor(px);
    double yy = py - std::floor(py);
    return ((1 - xx) * (1 - yy) * double(d[0]) + xx * (1 - yy) * double(d[1]) +
            (1 - xx) * yy * double(d[img.step]) + xx * yy * double(d[img.step + 1])) / 255.0;
}

// ref:449-480.  `heap` keeps the reference's two un-reserved std::vector<double>
// (ref:454,464-465) for the timed CPU baseline; the arithmetic and its order are the same
// either way.
template <bool HEAP>
double ncc(const Img8 &ref, const Img8 &curr, const V2 &pr, const V2 &pc) {
    constexpr int W = 3, AREA = 49;
    double mean_ref = 0, mean_curr = 0;
    double sr[AREA], sc[AREA];
    std::vector<double> vr, vc;
    int n = 0;
    for (int y = -W; y <= W; y++)
        for (int x = -W; x <= W; x++) {
            double value_ref = double(ref.data[size_t(int(y + pr.y)) * ref.step + size_t(int(x + pr.x))]) / 255.0;
            mean_ref += value_ref;
            double value_curr = bilinear(curr, pc.x + double(x), pc.y + double(y));
            mean_curr += value_curr;
            if (HEAP) { vr.push_back(value_ref); vc.push_back(value_curr); }
            else { sr[n] = value_ref; sc[n] = value_curr; }
            n++;
        }
    mean_ref /= AREA;
    mean_curr /= AREA;
    const double *a = HEAP ? vr.data() : sr;
    const double *b = HEAP ? vc.data() : sc;
    double numerator = 0, den1 = 0, den2 = 0;
    for (int i = 0; i < AREA; i++) {
        double m = (a[i] - mean_ref) * (b[i] - mean_curr);
        numerator += m;
        den1 += (a[i] - mean_ref) * (a[i] - mean_ref);
        den2 += (b[i] - mean_curr) * (b[i] - mean_curr);
    }
    return numerator / std::sqrt(den1 * den2 + 1e-10);
}

struct SearchOut {
    bool ok;
    V2 pt_curr, dir;
    double best_ncc;
    int n_eval;     // NCC() calls
    int n_steps;    // loop iterations of ref:432
    int best_step;  // iteration index of the winner (-1: none)
};

// ref:397-447
template <bool HEAP>
SearchOut epipolar_search(const Cam &c, const Img8 &ref, const Img8 &curr, const SE3 &T_C_R,
                          const V2 &pt_ref, double depth_mu, double depth_cov) {
    SearchOut o{};
    V3 f_ref = px2cam(c, pt_ref);
    normalize(f_ref);
    V3 P_ref = f_ref * depth_mu;
    V2 px_mean = cam2px(c, apply(T_C_R, P_ref));
    double d_min, d_max;
    if (c.inverse_depth) {  // ref:407-410
        const double inv_d_mu = 1.0 / depth_mu;
        const double inv_d_min = inv_d_mu - c.n_sigma * depth_cov, inv_d_max = inv_d_mu + c.n_sigma * depth_cov;
        d_min = 1.0 / inv_d_max;
        d_max = 1.0 / inv_d_min;
    } else {  // ref:412
        d_min = depth_mu - c.n_sigma * depth_cov;
        d_max = depth_mu + c.n_sigma * depth_cov;
    }
    if (d_min < c.min_depth) d_min = c.min_depth;  // ref:414
    V2 px_min = cam2px(c, apply(T_C_R, f_ref * d_min));
    V2 px_max = cam2px(c, apply(T_C_R, f_ref * d_max));
    V2 line{px_max.x - px_min.x, px_max.y - px_min.y};
    V2 dir = line;
    normalize(dir);
    double half_length = 0.5 * std::sqrt(line.x * line.x + line.y * line.y);
    if (half_length > c.max_half_len) half_length = c.max_half_len;  // ref:422

    double best_ncc = -1.0;
    V2 best_px{0, 0};  // uninitialised in the reference; only read after a win
    int it = 0;
    o.best_step = -1;
    for (double l = -half_length; l <= half_length; l += c.step, ++it) {  // ref:432
        V2 px{px_mean.x + l * dir.x, px_mean.y + l * dir.y};
        if (!inside(c, px)) continue;
        double v = ncc<HEAP>(ref, curr, pt_ref, px);
        o.n_eval++;
        if (v > best_ncc) { best_ncc = v; best_px = px; o.best_step = it; }
    }
    o.n_steps = it;
    o.best_ncc = best_ncc;
    o.dir = dir;
    o.pt_curr = best_px;
    o.ok = !(best_ncc < c.ncc_thresh);  // ref:443 `if (best_ncc < 0.85f) return false`
    return o;
}

// Eigen::ColPivHouseholderQR<Matrix2d>(A).solve(b), restated for the fixed 2x2 real case
// (computeInPlace + _solve_impl of Eigen 3.3/3.4).  A is row-major a[r][c].
void colpiv_qr_solve2(const double a_in[2][2], const double b_in[2], double x_out[2]) {
    const double eps = std::numeric_limits<double>::epsilon();
    double m[2][2] = {{a_in[0][0], a_in[0][1]}, {a_in[1][0], a_in[1][1]}};
    double norms_upd[2], norms_dir[2];
    for (int k = 0; k < 2; k++) {
        norms_dir[k] = std::sqrt(m[0][k] * m[0][k] + m[1][k] * m[1][k]);
        norms_upd[k] = norms_dir[k];
    }
    const double maxn = norms_upd[0] >= norms_upd[1] ? norms_upd[0] : norms_upd[1];
    const double threshold_helper = (maxn * eps) * (maxn * eps) / 2.0;
    const double norm_downdate_threshold = std::sqrt(eps);
    int nonzero_pivots = 2;
    int transp[2] = {0, 1};
    double hcoeff[2] = {0, 0};
    for (int k = 0; k < 2; k++) {
        int big = k;
        for (int j = k + 1; j < 2; j++)
            if (norms_upd[j] > norms_upd[big]) big = j;  // maxCoeff: first index on ties
        double big_sq = norms_upd[big] * norms_upd[big];
        if (nonzero_pivots == 2 && big_sq < threshold_helper * double(2 - k)) nonzero_pivots = k;
        transp[k] = big;
        if (k != big) {
            for (int r = 0; r < 2; r++) { double t = m[r][k]; m[r][k] = m[r][big]; m[r][big] = t; }
            double t = norms_upd[k]; norms_upd[k] = norms_upd[big]; norms_upd[big] = t;
            t = norms_dir[k]; norms_dir[k] = norms_dir[big]; norms_dir[big] = t;
        }
        // makeHouseholderInPlace on m[k..1][k]
        double c0 = m[k][k];
        double tail_sq = (k == 0) ? m[1][0] * m[1][0] : 0.0;
        double tau, beta;
        if (tail_sq <= std::numeric_limits<double>::min()) {
            tau = 0; beta = c0;
            if (k == 0) m[1][0] = 0;
        } else {
            beta = std::sqrt(c0 * c0 + tail_sq);
            if (c0 >= 0) beta = -beta;
            m[1][0] = m[1][0] / (c0 - beta);  // essential part (k == 0 only)
            tau = (beta - c0) / beta;
        }
        m[k][k] = beta;
        hcoeff[k] = tau;
        // apply H_k on the left to the trailing columns
        if (k == 0) {
            if (tau != 0) {
                double ess = m[1][0];
                double tmp = ess * m[1][1];
                tmp += m[0][1];
                m[0][1] -= tau * tmp;
                m[1][1] -= tau * ess * tmp;
            }
            // column-norm downdate for column 1
            if (norms_upd[1] != 0) {
                double temp = std::fabs(m[0][1]) / norms_upd[1];
                temp = (1.0 + temp) * (1.0 - temp);
                temp = temp < 0 ? 0 : temp;
                double r = norms_upd[1] / norms_dir[1];
                double temp2 = temp * (r * r);
                if (temp2 <= norm_downdate_threshold) {
                    norms_dir[1] = std::fabs(m[1][1]);
                    norms_upd[1] = norms_dir[1];
                } else {
                    norms_upd[1] *= std::sqrt(temp);
                }
            }
        }
    }
    // permutation indices from the transpositions
    int perm[2] = {0, 1};
    for (int k = 0; k < 2; k++) { int t = perm[k]; perm[k] = perm[transp[k]]; perm[transp[k]] = t; }

    x_out[0] = x_out[1] = 0;
    if (nonzero_pivots == 0) return;
    double c[2] = {b_in[0], b_in[1]};
    // c = Q^T b : apply H_0 (H_1 is the identity for a trailing 1-vector)
    if (hcoeff[0] != 0) {
        double ess = m[1][0];
        double tmp = ess * c[1];
        tmp += c[0];
        c[0] -= hcoeff[0] * tmp;
        c[1] -= hcoeff[0] * ess * tmp;
    }
    // back-substitution on the leading nonzero_pivots x nonzero_pivots upper triangle
    if (nonzero_pivots == 2) {
        c[1] = c[1] / m[1][1];
        c[0] = (c[0] - m[0][1] * c[1]) / m[0][0];
    } else {
        c[0] = c[0] / m[0][0];
    }
    for (int i = 0; i < nonzero_pivots; i++) x_out[perm[i]] = c[i];
}

struct FuseOut { double depth_est, d_cov2, mu_fuse, sigma_fuse2; };

// Diagnostic only (dmo_set_libm_perturbation, default 0 = off): moves the two acos results of ref:527,531 by up to
// +-n ulp, pseudo-randomly per call, to measure how far a libm that is not bit-identical to glibc (CUDA's acos / sin
// are specified to 1-2 ulp) can move the filter state.  Never enabled by tests that pin the oracle.
int g_perturb_ulp = 0;
inline double nudge(double v, double key) {
    uint64_t h;
    std::memcpy(&h, &key, 8);
    h ^= h >> 33; h *= 0xff51afd7ed558ccdULL; h ^= h >> 33; h *= 0xc4ceb9fe1a85ec53ULL; h ^= h >> 33;
    int k = int(h % uint64_t(2 * g_perturb_ulp + 1)) - g_perturb_ulp;
    for (; k > 0; --k) v = std::nextafter(v, 1e300);
    for (; k < 0; ++k) v = std::nextafter(v, -1e300);
    return v;
}

// ref:482-567
FuseOut update_depth_filter(const Cam &c, const V2 &pt_ref, const V2 &pt_curr, const SE3 &T_C_R,
                            const V2 &dir, double depth_val, double cov2_val) {
    SE3 T_R_C = inverse(T_C_R);
    V3 f_ref = px2cam(c, pt_ref);
    normalize(f_ref);
    V3 f_curr = px2cam(c, pt_curr);
    normalize(f_curr);
    const V3 t = T_R_C.t;
    const V3 f2 = rotate(T_R_C.q, f_curr);
    const double b[2] = {dot(t, f_ref), dot(t, f2)};
    double A[2][2];
    A[0][0] = dot(f_ref, f_ref);
    A[0][1] = -dot(f_ref, f2);
    A[1][0] = -A[0][1];
    A[1][1] = -dot(f2, f2);
    double ans[2];
    colpiv_qr_solve2(A, b, ans);
    const V3 xm = f_ref * ans[0];
    const V3 xn = t + f2 * ans[1];
    const V3 p_esti{(xm.x + xn.x) / 2.0, (xm.y + xn.y) / 2.0, (xm.z + xn.z) / 2.0};
    const double depth_estimation = norm(p_esti);

    const double t_norm = norm(t);
    double alpha = std::acos(dot(f_ref, t) / t_norm);
    V3 f_curr_prime = px2cam(c, V2{pt_curr.x + dir.x, pt_curr.y + dir.y});
    normalize(f_curr_prime);
    const V3 mt{-t.x, -t.y, -t.z};
    double beta_prime = std::acos(dot(f_curr_prime, mt) / t_norm);  // not rotated into the ref frame (ref:529-531)
    if (g_perturb_ulp) {  // diagnostic (tools/libm_sensitivity.py): what a libm that differs by a few ulp does to the filter
        alpha = nudge(alpha, pt_ref.x * 31.0 + pt_ref.y * 17.0 + pt_curr.x);
        beta_prime = nudge(beta_prime, pt_ref.x * 13.0 + pt_ref.y * 29.0 + pt_curr.y);
    }
    const double gamma = M_PI - alpha - beta_prime;
    const double p_prime_norm = t_norm * std::sin(beta_prime) / std::sin(gamma);
    const double d_cov = c.inverse_depth ? (1.0 / p_prime_norm - 1.0 / depth_estimation)
                                         : (p_prime_norm - depth_estimation);
    const double d_cov2 = d_cov * d_cov;

    const double mu = c.inverse_depth ? 1.0 / depth_val : depth_val;
    const double sigma2 = cov2_val;
    const double meas = c.inverse_depth ? 1.0 / depth_estimation : depth_estimation;
    // ref:552 writes `sigma2 * 1.0 / depth_estimation` = (sigma2*1.0)/depth_estimation in the inverse arm
    const double mu_fuse = c.inverse_depth
                               ? (d_cov2 * mu + sigma2 * 1.0 / depth_estimation) / (sigma2 + d_cov2 + 1e-10)
                               : (d_cov2 * mu + sigma2 * meas) / (sigma2 + d_cov2 + 1e-10);
    const double sigma_fuse2 = (sigma2 * d_cov2) / (sigma2 + d_cov2 + 1e-10);
    return {depth_estimation, d_cov2, c.inverse_depth ? 1.0 / mu_fuse : mu_fuse, sigma_fuse2};
}

template <bool HEAP>
void update_rows(const Cam &c, const Img8 &ref, const Img8 &curr, const SE3 &T,
                 double *depth, size_t dstep, double *cov2, size_t cstep,
                 int row_begin, int row_end, int row_stride,
                 dmf_counters *counters, uint8_t *flags, size_t fstep,
                 float *dbg_ncc, int32_t *dbg_n, size_t dbg_w, int32_t *dbg_k = nullptr, double *dbg_ncc64 = nullptr) {
    int y0 = row_begin < c.border ? c.border : row_begin;
    int y1 = row_end > c.height - c.border ? c.height - c.border : row_end;
    if (row_stride < 1) row_stride = 1;
    // align y0 to the stride grid anchored at row_begin
    if (row_begin < y0) { int k = (y0 - row_begin + row_stride - 1) / row_stride; y0 = row_begin + k * row_stride; }
    long long n_rows = y1 > y0 ? (y1 - y0 + row_stride - 1) / row_stride : 0;
    unsigned long long interior = 0, active = 0, evals = 0, accepted = 0;
    // The reference parallelises over rows (ref:356).  Rows are additionally cut into column blocks here so
    // that a ROW SUBSET (the bounded CPU-baseline sample of bench.py, fewer rows than cores) still keeps every
    // core busy; pixels are independent, so the arithmetic per pixel is unchanged.
    const int XB = 64;
    const long long n_xb = (c.width - 2 * c.border + XB - 1) / XB;
#pragma omp parallel for collapse(2) schedule(dynamic, 1) reduction(+ : interior, active, evals, accepted)
    for (long long r = 0; r < n_rows; r++) {
      for (long long xb = 0; xb < n_xb; xb++) {
        int y = y0 + int(r) * row_stride;
        double *drow = reinterpret_cast<double *>(reinterpret_cast<char *>(depth) + size_t(y) * dstep);
        double *crow = reinterpret_cast<double *>(reinterpret_cast<char *>(cov2) + size_t(y) * cstep);
        const int xa = c.border + int(xb) * XB;
        const int xe = (xa + XB < c.width - c.border) ? xa + XB : c.width - c.border;
        for (int x = xa; x < xe; x++) {  // ref:363
            interior++;
            uint8_t fl = 0;
            if (!(crow[x] < c.min_cov || crow[x] > c.max_cov)) {  // ref:366 (NaN passes)
                fl |= 1;
                active++;
                SearchOut s = epipolar_search<HEAP>(c, ref, curr, T, V2{double(x), double(y)}, drow[x],
                                                    std::sqrt(crow[x]));
                evals += (unsigned long long)s.n_eval;
                if (dbg_ncc) dbg_ncc[size_t(y) * dbg_w + x] = (float)s.best_ncc;
                if (dbg_n) dbg_n[size_t(y) * dbg_w + x] = s.n_eval;
                if (dbg_k) dbg_k[size_t(y) * dbg_w + x] = (s.n_steps << 16) | (s.best_step < 0 ? 0xFFFF : s.best_step);
                if (dbg_ncc64) dbg_ncc64[size_t(y) * dbg_w + x] = s.best_ncc;
                if (s.ok) {
                    fl |= 2;
                    accepted++;
                    FuseOut f = update_depth_filter(c, V2{double(x), double(y)}, s.pt_curr, T, s.dir, drow[x], crow[x]);
                    drow[x] = f.mu_fuse;        // ref:562
                    crow[x] = f.sigma_fuse2;    // ref:564
                }
            }
            if (flags) flags[size_t(y) * fstep + x] = fl;
        }
      }
    }
    if (counters) {
        counters->frames += 1;
        counters->interior += interior;
        counters->active += active;
        counters->ncc_evals += evals;
        counters->accepted += accepted;
    }
}

inline SE3 make_se3(const double q[4], const double t[3]) { return {Quat{q[0], q[1], q[2], q[3]}, V3{t[0], t[1], t[2]}}; }

}  // namespace

extern "C" {

// Reference constants ref:72-89 (same contract as dmf_default_params in include/dmf.h;
// restated here so the oracle library stands alone).
int dmo_default_params(dmf_params *p, int width, int height, int inverse_depth) {
    if (!p || width <= 0 || height <= 0) return -1;
    std::memset(p, 0, sizeof(*p));
    p->width = width; p->height = height; p->border = 20; p->ncc_half = 3;
    if (width == 640 && height == 480) {
        p->fx = 481.2f; p->fy = -480.0f; p->cx = 319.5f; p->cy = 239.5f;
    } else {
        const double s = double(width) / 640.0;
        p->fx = double(481.2f) * s; p->fy = -480.0 * s;
        p->cx = 0.5 * (width - 1); p->cy = 0.5 * (height - 1);
    }
    p->step = 0.7; p->max_half_len = 100; p->min_depth = 0.1; p->n_sigma = 3;
    p->ncc_thresh = 0.85f;
    if (inverse_depth) { p->min_cov = 0.0001; p->max_cov = 1; }
    else { const double good_error = 0.01; p->min_cov = good_error * good_error; p->max_cov = 10; }
    p->inverse_depth = inverse_depth ? 1 : 0;
    return 0;
}

// One update() (ref:355-393) over interior rows {row_begin + k*row_stride} ∩ [border, H-border).
// heap != 0 keeps the reference's per-NCC std::vector allocations (timed baseline).
// flags (optional, full-image, fstep bytes/row): bit0 = passed gate ref:366, bit1 = accepted ref:443.
// dbg_ncc / dbg_n (optional, full-image W-strided): best NCC and number of NCC calls per pixel.
int dmo_update(const dmf_params *p, const uint8_t *ref, size_t ref_step, const uint8_t *curr, size_t curr_step,
               const double q_xyzw[4], const double t_xyz[3], double *depth, size_t depth_step,
               double *cov2, size_t cov2_step, int row_begin, int row_end, int row_stride, int heap,
               dmf_counters *counters, uint8_t *flags, size_t flags_step, float *dbg_ncc, int32_t *dbg_n) {
    if (!p || !ref || !curr || !depth || !cov2 || !q_xyzw || !t_xyz || p->ncc_half != 3) return -1;
    Cam c = make_cam(*p);
    Img8 r{ref, ref_step}, cu{curr, curr_step};
    SE3 T = make_se3(q_xyzw, t_xyz);
    if (heap)
        update_rows<true>(c, r, cu, T, depth, depth_step, cov2, cov2_step, row_begin, row_end, row_stride, counters,
                          flags, flags_step, dbg_ncc, dbg_n, size_t(p->width));
    else
        update_rows<false>(c, r, cu, T, depth, depth_step, cov2, cov2_step, row_begin, row_end, row_stride, counters,
                           flags, flags_step, dbg_ncc, dbg_n, size_t(p->width));
    return 0;
}

// dmo_update with two more optional full-image planes (tools/parity_diag.py): dbg_k = (trip count of the loop
// ref:432 << 16) | index of the winning iteration (0xFFFF: none) — the layout of dmf_download_debug — and the
// best NCC as a double.
int dmo_update_ex(const dmf_params *p, const uint8_t *ref, size_t ref_step, const uint8_t *curr, size_t curr_step,
                  const double q_xyzw[4], const double t_xyz[3], double *depth, size_t depth_step,
                  double *cov2, size_t cov2_step, int row_begin, int row_end, int row_stride,
                  dmf_counters *counters, uint8_t *flags, size_t flags_step, int32_t *dbg_k, double *dbg_ncc64) {
    if (!p || !ref || !curr || !depth || !cov2 || !q_xyzw || !t_xyz || p->ncc_half != 3) return -1;
    Cam c = make_cam(*p);
    Img8 r{ref, ref_step}, cu{curr, curr_step};
    update_rows<false>(c, r, cu, make_se3(q_xyzw, t_xyz), depth, depth_step, cov2, cov2_step, row_begin, row_end, row_stride,
                       counters, flags, flags_step, nullptr, nullptr, size_t(p->width), dbg_k, dbg_ncc64);
    return 0;
}

// Unit-level entry points (per-function fixtures).
double dmo_bilinear(const uint8_t *img, size_t step, double x, double y) { return bilinear(Img8{img, step}, x, y); }

double dmo_ncc(const uint8_t *ref, size_t ref_step, const uint8_t *curr, size_t curr_step, double rx, double ry,
               double cx_, double cy_) {
    return ncc<false>(Img8{ref, ref_step}, Img8{curr, curr_step}, V2{rx, ry}, V2{cx_, cy_});
}

// out[0..7] = ok, pt_curr.x, pt_curr.y, dir.x, dir.y, best_ncc, n_eval, n_steps ; out[8] = best_step
int dmo_epipolar_search(const dmf_params *p, const uint8_t *ref, size_t ref_step, const uint8_t *curr,
                        size_t curr_step, const double q[4], const double t[3], double x, double y, double mu,
                        double sigma, double out[9]) {
    Cam c = make_cam(*p);
    SearchOut s = epipolar_search<false>(c, Img8{ref, ref_step}, Img8{curr, curr_step}, make_se3(q, t), V2{x, y}, mu, sigma);
    out[0] = s.ok; out[1] = s.pt_curr.x; out[2] = s.pt_curr.y; out[3] = s.dir.x; out[4] = s.dir.y;
    out[5] = s.best_ncc; out[6] = s.n_eval; out[7] = s.n_steps; out[8] = s.best_step;
    return 0;
}

// out = depth_est, d_cov2, fused depth (as stored), fused cov2
int dmo_update_depth_filter(const dmf_params *p, const double q[4], const double t[3], double rx, double ry,
                            double cx_, double cy_, double dirx, double diry, double depth_val, double cov2_val,
                            double out[4]) {
    Cam c = make_cam(*p);
    FuseOut f = update_depth_filter(c, V2{rx, ry}, V2{cx_, cy_}, make_se3(q, t), V2{dirx, diry}, depth_val, cov2_val);
    out[0] = f.depth_est; out[1] = f.d_cov2; out[2] = f.mu_fuse; out[3] = f.sigma_fuse2;
    return 0;
}

void dmo_qr_solve2(const double a[4], const double b[2], double x[2]) {
    double A[2][2] = {{a[0], a[1]}, {a[2], a[3]}};
    colpiv_qr_solve2(A, b, x);
}

// T_C_R = T_WC(curr)^-1 * T_WC(ref)   (ref:289-290); inputs as read by ref:333-335
// (quaternion normalised by the SE3d constructor).  Layout q = (x,y,z,w).
void dmo_compose_T_C_R(const double q_ref[4], const double t_ref[3], const double q_cur[4], const double t_cur[3],
                       double q_out[4], double t_out[3]) {
    SE3 Tr = make_se3(q_ref, t_ref), Tc = make_se3(q_cur, t_cur);
    Tr.q = normalized(Tr.q);
    Tc.q = normalized(Tc.q);
    SE3 T = compose(inverse(Tc), Tr);
    q_out[0] = T.q.x; q_out[1] = T.q.y; q_out[2] = T.q.z; q_out[3] = T.q.w;
    t_out[0] = T.t.x; t_out[1] = T.t.y; t_out[2] = T.t.z;
}

void dmo_transform_point(const double q[4], const double t[3], const double p[3], double out[3]) {
    V3 r = apply(make_se3(q, t), V3{p[0], p[1], p[2]});
    out[0] = r.x; out[1] = r.y; out[2] = r.z;
}

// evaludateDepth ref:569-590 over rows [row_begin,row_end) ∩ interior: returns sum of squared
// errors and the count; RMS = sqrt(sum/count) as printed at ref:589.
int dmo_evaluate_depth(const dmf_params *p, const double *truth, size_t tstep, const double *est, size_t estep,
                       const double *var, size_t vstep, double max_variance, int row_begin, int row_end,
                       double *sum_sq, uint64_t *count) {
    int y0 = row_begin < p->border ? p->border : row_begin;
    int y1 = row_end > p->height - p->border ? p->height - p->border : row_end;
    double s = 0; uint64_t n = 0;
    for (int y = y0; y < y1; y++) {
        const double *tr = reinterpret_cast<const double *>(reinterpret_cast<const char *>(truth) + size_t(y) * tstep);
        const double *er = reinterpret_cast<const double *>(reinterpret_cast<const char *>(est) + size_t(y) * estep);
        const double *vr = reinterpret_cast<const double *>(reinterpret_cast<const char *>(var) + size_t(y) * vstep);
        for (int x = p->border; x < p->width - p->border; x++) {
            if (vr[x] >= max_variance) continue;  // ref:579
            double e = tr[x] - er[x];
            s += e * e;
            n++;
        }
    }
    *sum_sq = s; *count = n;
    return 0;
}

// getMaskFromVariance ref:199-204: threshold(THRESH_BINARY_INV, 255) then convertTo(CV_8U):
// mask = var > max_variance ? 0 : 255 (NaN compares false -> 255), whole image.
int dmo_variance_mask(int width, int height, const double *var, size_t vstep, double max_variance, uint8_t *mask,
                      size_t mstep) {
    for (int y = 0; y < height; y++) {
        const double *vr = reinterpret_cast<const double *>(reinterpret_cast<const char *>(var) + size_t(y) * vstep);
        for (int x = 0; x < width; x++) mask[size_t(y) * mstep + x] = vr[x] > max_variance ? 0 : 255;
    }
    return 0;
}

// getPointCloudFromImageAndDistance, utils/pointcloud/pointcloud_from_image_depth.h:42-89, with
// T = identity (ref:283) and a 3-channel BGR colour image.  xyz: 3 doubles per point (the
// reference narrows to float in PointXYZRGB; callers compare after the same cast), rgb: 3 bytes.
// Returns the number of points written (row-major scan order).
long long dmo_point_cloud(const dmf_params *p, const uint8_t *color, size_t color_step, int channels,
                          const double *dist, size_t dstep, const uint8_t *mask, size_t mstep, float *xyz,
                          uint8_t *rgb, long long capacity) {
    long long n = 0;
    for (int v = p->border; v < p->height - p->border; v++) {
        const double *dr = reinterpret_cast<const double *>(reinterpret_cast<const char *>(dist) + size_t(v) * dstep);
        for (int u = p->border; u < p->width - p->border; u++) {
            const double d = dr[u];
            const uint8_t valid = mask ? mask[size_t(v) * mstep + u] : 1;
            if (d == 0 || valid == 0) continue;
            V3 pt{(u - p->cx) / p->fx, (v - p->cy) / p->fy, 1.0};
            normalize(pt);
            pt = pt * d;
            if (n < capacity) {
                xyz[3 * n + 0] = (float)pt.x; xyz[3 * n + 1] = (float)pt.y; xyz[3 * n + 2] = (float)pt.z;
                const uint8_t *c = &color[size_t(v) * color_step + size_t(u) * channels];
                rgb[3 * n + 0] = channels >= 3 ? c[2] : c[0];  // r
                rgb[3 * n + 1] = channels >= 3 ? c[1] : c[0];  // g
                rgb[3 * n + 2] = c[0];                          // b
            }
            n++;
        }
    }
    return n;
}

int dmo_max_threads(void);
void dmo_set_libm_perturbation(int ulp) { g_perturb_ulp = ulp < 0 ? 0 : ulp; }

}  // extern "C"

#ifdef _OPENMP
#include <omp.h>
extern "C" int dmo_max_threads(void) { return omp_get_max_threads(); }
extern "C" void dmo_set_threads(int n) { if (n > 0) omp_set_num_threads(n); }
#else
extern "C" int dmo_max_threads(void) { return 1; }
extern "C" void dmo_set_threads(int) {}
#endif
