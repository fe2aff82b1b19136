"""CPU restatement of the FP64 geometry of the CUDA path (search_geometry / fuse_slot in
slamplay_b200/csrc/dmf_kernels.cuh) checked against the oracle (reference epipolarSearch ref:397-447 and
updateDepthFilter ref:482-567).

The kernels do not call what the reference calls: unit rays by one reciprocal square root, projections by one
reciprocal per point, a Cramer solve instead of ColPivHouseholderQR, and the one-pixel uncertainty without acos / sin
(sin(acos c) = sqrt(1 - c^2), sin(pi - a - b) = sin(a + b)).  This test pins those identities and their error level
on the golden unit vectors of the compiled reference, without a GPU."""
import ctypes as C
from pathlib import Path

import numpy as np

import oracle

G = Path(__file__).resolve().parent / "golden"


def qrot(q, v):  # Eigen Quaternion::_transformVector, as Sophus SE3 * point uses it
    qv = np.array(q[:3])
    uv = 2.0 * np.cross(qv, v)
    return v + q[3] * uv + np.cross(qv, uv)


def se3_inverse(q, t):  # Sophus SE3::inverse(): conjugate (normalised), t' = R^-1 * (-t)
    c = np.array([-q[0], -q[1], -q[2], q[3]])
    c = c / np.sqrt((c[0] * c[0] + c[2] * c[2]) + (c[1] * c[1] + c[3] * c[3]))
    return c, qrot(c, -np.array(t))


def unit_ray(p, u, v):  # normalize(px2cam(u, v)) with one rsqrt
    X, Y = (u - p.cx) * (1.0 / p.fx), (v - p.cy) * (1.0 / p.fy)
    r = 1.0 / np.sqrt(X * X + Y * Y + 1.0)
    return np.array([X * r, Y * r, r])


def search_geometry(p, q, t, x, y, mu, sigma):
    """px_mean, unit direction and half length of the epipolar segment, the kernel's way (ref:402-422)."""
    Rf = qrot(q, unit_ray(p, x, y))
    d_min, d_max = max(mu - p.n_sigma * sigma, p.min_depth), mu + p.n_sigma * sigma

    def proj(d):
        rz = 1.0 / (Rf[2] * d + t[2])
        return np.array([(Rf[0] * d + t[0]) * p.fx * rz + p.cx, (Rf[1] * d + t[1]) * p.fy * rz + p.cy])

    pm, p0, p1 = proj(mu), proj(d_min), proj(d_max)
    line = p1 - p0
    length = np.sqrt(line @ line)
    direction = line / length if length > 0 else line
    return pm, direction, min(0.5 * length, p.max_half_len)


def fuse(p, q, t, x, y, cxp, cyp, ex, ey, mu, c2):
    """depth_est, d_cov2, fused depth, fused cov2, the kernel's way (ref:482-567, non-inverse-depth variant)."""
    qi, ti = se3_inverse(q, t)
    f_ref, f_curr = unit_ray(p, x, y), unit_ray(p, cxp, cyp)
    f2 = qrot(qi, f_curr)
    b0, b1 = ti @ f_ref, ti @ f2
    a00, a01, a11 = f_ref @ f_ref, -(f_ref @ f2), -(f2 @ f2)
    a10 = -a01
    rdet = 1.0 / (a00 * a11 - a01 * a10)
    ans0, ans1 = (b0 * a11 - a01 * b1) * rdet, (a00 * b1 - a10 * b0) * rdet
    pe = 0.5 * (ans0 * f_ref + (ti + ans1 * f2))
    depth_est = np.sqrt(pe @ pe)
    t_norm = np.sqrt(ti @ ti)
    ca = (f_ref @ ti) / t_norm
    cb = -(unit_ray(p, cxp + ex, cyp + ey) @ ti) / t_norm
    sa, sb = np.sqrt(1.0 - ca * ca), np.sqrt(1.0 - cb * cb)
    p_prime = t_norm * sb / (sa * cb + ca * sb)
    d_cov2 = (p_prime - depth_est) ** 2
    rden = 1.0 / (c2 + d_cov2 + 1e-10)
    return depth_est, d_cov2, (d_cov2 * mu + c2 * depth_est) * rden, (c2 * d_cov2) * rden


def _golden():
    u = np.load(G / "remode640_ref_units.npz")
    p = oracle.default_params(640, 480)
    return u, p, u["pose"][:4], u["pose"][4:]


def test_fusion_without_qr_and_transcendentals_matches_the_reference_arithmetic():
    u, p, q, t = _golden()
    L = oracle.lib()
    qa, ta = (C.c_double * 4)(*q), (C.c_double * 3)(*t)
    worst = np.zeros(4)
    n = 0
    for i in range(len(u["rx"])):
        out = (C.c_double * 4)()
        L.dmo_update_depth_filter(C.byref(p), qa, ta, u["rx"][i], u["ry"][i], u["cx"][i], u["cy"][i], u["dirs"][i, 0],
                                  u["dirs"][i, 1], u["dval"][i], u["cval"][i], out)
        want = np.array(out[:])
        if not np.all(np.isfinite(want)):
            continue
        got = np.array(fuse(p, q, t, u["rx"][i], u["ry"][i], u["cx"][i], u["cy"][i], u["dirs"][i, 0], u["dirs"][i, 1],
                            u["dval"][i], u["cval"][i]))
        worst = np.maximum(worst, np.abs(got - want) / np.maximum(np.abs(want), 1e-300))
        n += 1
    assert n > 200
    # depth_est / fused depth: O(cond * eps); the variances square a difference of two nearly equal lengths
    assert worst[0] < 1e-9 and worst[2] < 1e-9, worst
    assert worst[1] < 1e-6 and worst[3] < 1e-6, worst


def test_search_segment_matches_the_reference_arithmetic():
    u, p, q, t = _golden()
    L = oracle.lib()
    qa, ta = (C.c_double * 4)(*q), (C.c_double * 3)(*t)
    ref = np.zeros((480, 640), np.uint8)  # images only feed the NCC; direction and step count do not depend on them
    worst_dir, n = 0.0, 0
    for i in range(len(u["rx"])):
        out = (C.c_double * 9)()
        L.dmo_epipolar_search(C.byref(p), ref.ctypes.data, 640, ref.ctypes.data, 640, qa, ta, u["rx"][i], u["ry"][i],
                              u["mu"][i], u["sigma"][i], out)
        pm, direction, half = search_geometry(p, q, t, u["rx"][i], u["ry"][i], u["mu"][i], u["sigma"][i])
        worst_dir = max(worst_dir, abs(direction[0] - out[3]), abs(direction[1] - out[4]))
        # trip count of `for (l = -half; l <= half; l += 0.7)` (ref:432) from the closed form the kernels use
        k = int(2.0 * half / p.step) + 1
        while k > 0 and (p.step * (k - 1) - half) > half:
            k -= 1
        while (p.step * k - half) <= half:
            k += 1
        assert k == int(out[7]), (k, out[7], half)
        n += 1
    assert n == len(u["rx"]) and worst_dir < 1e-10
