"""Parity metrics of SURVEY.md §8d (P1 depth agreement, P2 decision agreement) shared by the tests."""
import numpy as np

DEPTH_RTOL = 1e-3        # north_star: converged depth within 1e-3 relative per pixel
MIN_DEPTH_AGREE = 0.95   # north_star: at least 95 % of pixels agree
MAX_DECISION_MISMATCH = 0.005  # north_star: accept/reject + covariance gates differ on at most 0.5 % of pixels


def interior(p, rows=None):
    b = p.border
    ys = np.arange(b, p.height - b) if rows is None else np.array([y for y in rows if b <= y < p.height - b])
    return ys, slice(b, p.width - b)


def state_class(p, cov2):
    """{0 converged, 1 diverged, 2 active, 3 NaN} per the gates of ref:366."""
    return np.where(np.isnan(cov2), 3, np.where(cov2 < p.min_cov, 0, np.where(cov2 > p.max_cov, 1, 2)))


def depth_agreement(p, d_test, d_ref, rows=None, rtol=DEPTH_RTOL):
    ys, xs = interior(p, rows)
    a, b = d_test[ys][:, xs], d_ref[ys][:, xs]
    both_nan = np.isnan(a) & np.isnan(b)
    return float(((np.abs(a - b) <= rtol * np.abs(b)) | both_nan).mean())


def class_mismatch(p, c_test, c_ref, rows=None):
    ys, xs = interior(p, rows)
    return float((state_class(p, c_test[ys][:, xs]) != state_class(p, c_ref[ys][:, xs])).mean())


def flag_mismatch(p, f_test, f_ref, rows=None):
    ys, xs = interior(p, rows)
    return float((f_test[ys][:, xs] != f_ref[ys][:, xs]).mean())


def synthetic_state(h=480, w=640):
    """Deterministic state / truth maps from integer hashes (no RNG stream to drift): depth in [1.5, 3.5), cov2 log-uniform
    over [1e-5, 10) with a few NaN and zero-depth pixels, truth = depth + a small offset."""
    y, x = np.mgrid[0:h, 0:w].astype(np.uint64)
    hsh = lambda a, b: ((x * np.uint64(a)) ^ (y * np.uint64(b))) % np.uint64(1000003)
    u1 = hsh(73856093, 19349663).astype(np.float64) / 1000003.0
    u2 = hsh(83492791, 49979687).astype(np.float64) / 1000003.0
    u3 = hsh(2654435761, 40503).astype(np.float64) / 1000003.0
    depth = 1.5 + 2.0 * u1
    cov2 = 10.0 ** (-5.0 + 6.0 * u2)
    truth = depth + 0.05 * (u3 - 0.5)
    depth[::53, ::47] = 0.0      # "not measured" (pointcloud_from_image_depth.h:66)
    cov2[::41, ::59] = np.nan    # NaN variance: counted by evaludateDepth? (NaN >= t is false -> counted), mask 255
    return depth, cov2, truth
