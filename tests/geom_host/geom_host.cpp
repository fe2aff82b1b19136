// Host build of slamplay_b200/csrc/dmf_geometry.h (the FP64 geometry the CUDA kernels run, in the reference's operation
// order) for the CPU test-suite: tests/test_geometry_formulation.py compares it bit for bit with the oracle.
// Built by that test with g++ -O2 -ffp-contract=off (no FMA contraction, like the oracle and the compiled reference).
#include "../../slamplay_b200/csrc/dmf_geometry.h"

using namespace dmf_geom;

extern "C" {

// out: pm.x pm.y dir.x dir.y half n
void gh_search(const double cam[4], const double q[4], const double t[3], double x, double y, double mu, double sigma,
               double n_sigma, double min_depth, double max_half_len, double step, int inverse_depth, double out[6]) {
    const Camera c{cam[0], cam[1], cam[2], cam[3]};
    const V3 f = unit_ray(c, x, y);
    const Segment s = search_segment(c, q, t, f, mu, sigma, n_sigma, min_depth, max_half_len, inverse_depth != 0);
    out[0] = s.pm.x; out[1] = s.pm.y; out[2] = s.dir.x; out[3] = s.dir.y; out[4] = s.half;
    out[5] = (double)trip_count(s.half, step);
}

// position of sample k of the segment (accumulated l)
void gh_sample(const double pm[2], const double dir[2], double half, double step, int k, double out[2]) {
    const V2 p = sample_pos(V2{pm[0], pm[1]}, V2{dir[0], dir[1]}, sample_l_acc(half, step, k));
    out[0] = p.x; out[1] = p.y;
}

int gh_trip_count(double half, double step) { return trip_count(half, step); }

// out: depth_est d_cov2 mu sigma2
void gh_fuse(const double cam[4], const double qi[4], const double ti[3], double t_norm, double x, double y, double cx, double cy,
             double dx, double dy, double mu, double sigma2, int inverse_depth, double out[4]) {
    const Camera c{cam[0], cam[1], cam[2], cam[3]};
    const Fused f = fuse(c, qi, ti, t_norm, unit_ray(c, x, y), V2{cx, cy}, V2{dx, dy}, mu, sigma2, inverse_depth != 0);
    out[0] = f.depth_est; out[1] = f.d_cov2; out[2] = f.mu; out[3] = f.sigma2;
}

void gh_qr_solve2(const double a[4], const double b[2], double x[2]) { colpiv_qr_solve2(a[0], a[1], a[2], a[3], b[0], b[1], x[0], x[1]); }

}  // extern "C"
