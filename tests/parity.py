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
