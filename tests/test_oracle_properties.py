"""Self-checks of the oracle that need no reference build: analytic cases and the edge
semantics the reference has and the CUDA path must keep."""
import ctypes as C

import numpy as np
import pytest

import oracle
from slamplay_b200.se3 import SE3, relative_pose
from slamplay_b200.synth import make_params, make_sequence


def _d(*v):
    return (C.c_double * len(v))(*v)


def test_default_params_are_the_reference_constants():
    p = oracle.default_params(640, 480)
    assert (p.width, p.height, p.border, p.ncc_half) == (640, 480, 20, 3)                   # ref:72-74,79
    assert p.fx == float(np.float32(481.2)) and p.fy == -480.0 and p.cx == 319.5 and p.cy == 239.5  # ref:75-78
    assert p.fx == 481.20001220703125
    assert p.min_cov == 0.01 * 0.01 and p.max_cov == 10 and p.ncc_thresh == float(np.float32(0.85))  # ref:85-87,443
    q = oracle.default_params(640, 480, True)
    assert (q.min_cov, q.max_cov, q.inverse_depth) == (0.0001, 1.0, 1)                      # ref:82-83
    m = make_params(640, 480)
    assert bytes(m) == bytes(p), "python make_params and the C default must agree"


def test_bilinear_matches_formula():
    rng = np.random.default_rng(0)
    img = rng.integers(0, 256, (32, 48), dtype=np.uint8)
    L = oracle.lib()
    for _ in range(100):
        x, y = rng.uniform(1, 46), rng.uniform(1, 30)
        ix, iy = int(x), int(y)
        xx, yy = x - np.floor(x), y - np.floor(y)
        exp = ((1 - xx) * (1 - yy) * float(img[iy, ix]) + xx * (1 - yy) * float(img[iy, ix + 1])
               + (1 - xx) * yy * float(img[iy + 1, ix]) + xx * yy * float(img[iy + 1, ix + 1])) / 255.0
        assert L.dmo_bilinear(img.ctypes.data, img.strides[0], x, y) == exp  # ref:165-174


def test_ncc_identity_and_affine_invariance():
    rng = np.random.default_rng(1)
    a = rng.integers(40, 200, (64, 64), dtype=np.uint8)
    L = oracle.lib()
    v = L.dmo_ncc(a.ctypes.data, 64, a.ctypes.data, 64, 30.0, 30.0, 30.0, 30.0)
    assert abs(v - 1.0) < 1e-9
    b = (a.astype(np.int32) // 2 + 17).astype(np.uint8)  # gain + offset: ZNCC stays ~1 (quantisation aside)
    v2 = L.dmo_ncc(a.ctypes.data, 64, b.ctypes.data, 64, 30.0, 30.0, 30.0, 30.0)
    assert v2 > 0.995
    flat = np.full((64, 64), 90, np.uint8)
    assert abs(L.dmo_ncc(a.ctypes.data, 64, flat.ctypes.data, 64, 30.0, 30.0, 30.0, 30.0)) < 1e-12  # ~0 / sqrt(1e-10)


def test_qr_solve_regular_singular_and_zero():
    L = oracle.lib()
    rng = np.random.default_rng(2)
    for _ in range(200):
        A = rng.normal(size=(2, 2))
        b = rng.normal(size=2)
        x = (C.c_double * 2)()
        L.dmo_qr_solve2(_d(*A.ravel()), _d(*b), x)
        assert np.allclose(A @ np.array(list(x)), b, rtol=1e-9, atol=1e-9)
    x = (C.c_double * 2)()
    # exactly-zero matrix: Eigen's pivot threshold is 0 there, so no pivot is discarded and the
    # back-substitution divides by zero (never reached on the path: A(0,0) = f_ref.f_ref = 1, ref:507)
    L.dmo_qr_solve2(_d(0, 0, 0, 0), _d(1, 2), x)
    assert not np.isfinite(list(x)).all()
    L.dmo_qr_solve2(_d(1, 2, 2, 4), _d(1, 2), x)  # rank 1 and consistent: finite solution of the system
    assert np.allclose(np.array([[1, 2], [2, 4]]) @ np.array(list(x)), [1, 2])


def test_inside_is_asymmetric():
    """ref:222-224: x + border < width but y + border <= height."""
    p = oracle.default_params(640, 480)
    img = np.random.default_rng(3).integers(0, 256, (480, 640), dtype=np.uint8)
    L = oracle.lib()
    out = (C.c_double * 9)()
    # identity pose: zero-length segment -> exactly one sample at the pixel itself
    q, t = _d(0, 0, 0, 1), _d(0, 0, 0)
    L.dmo_epipolar_search(C.byref(p), img.ctypes.data, 640, img.ctypes.data, 640, q, t, 300.0, 200.0, 2.0, 0.5, out)
    assert out[7] == 1 and out[6] == 1 and out[0] == 1.0 and (out[3], out[4]) == (0.0, 0.0)
    # a sample at y == height - border is inside, at x == width - border it is not
    pp = oracle.default_params(640, 480)
    L.dmo_epipolar_search(C.byref(pp), img.ctypes.data, 640, img.ctypes.data, 640, q, t, 300.0, 460.0, 2.0, 0.5, out)
    assert out[6] == 1  # y + 20 <= 480
    L.dmo_epipolar_search(C.byref(pp), img.ctypes.data, 640, img.ctypes.data, 640, q, t, 620.0, 200.0, 2.0, 0.5, out)
    assert out[6] == 0  # x + 20 < 640 fails


def test_gates_and_nan_semantics(seq640):
    """ref:366: converged / diverged pixels are skipped; NaN passes the gate but yields no sample."""
    seq, frames = seq640
    h, w = seq.shape
    depth = np.full((h, w), 2.0)
    cov2 = np.full((h, w), 0.5)
    cov2[100, 100] = 0.5e-4   # converged
    cov2[100, 101] = 10.5     # diverged
    cov2[100, 102] = np.nan   # NaN variance
    depth[100, 103] = np.nan  # NaN depth
    flags = np.zeros((h, w), np.uint8)
    cnt = oracle.Counters()
    T = seq.T_C_R(2)
    d0, c0 = depth.copy(), cov2.copy()
    oracle.update(seq.params, frames[0], frames[2], T.q, T.t, depth, cov2, rows=(100, 101), flags=flags, counters=cnt)
    assert flags[100, 100] == 0 and flags[100, 101] == 0
    assert flags[100, 102] == 1 and flags[100, 103] == 1  # active, not accepted
    for x in (100, 101, 102, 103):
        assert np.array_equal(depth[100, x], d0[100, x], equal_nan=True) and np.array_equal(cov2[100, x], c0[100, x], equal_nan=True)
    assert cnt.interior == w - 40 and cnt.active == w - 40 - 2
    # rows outside [border, H-border) and columns in the border are never touched
    assert np.array_equal(depth[:100], d0[:100], equal_nan=True) and np.array_equal(depth[101:], d0[101:], equal_nan=True)


def test_zero_baseline_poisons_with_nan(seq640):
    """|t| = 0 => acos(0/0) (ref:527): every accepted pixel becomes NaN, as in the reference."""
    seq, frames = seq640
    h, w = seq.shape
    depth = np.full((h, w), 2.0)
    cov2 = np.full((h, w), 0.5)
    flags = np.zeros((h, w), np.uint8)
    oracle.update(seq.params, frames[0], frames[0], (0, 0, 0, 1), (0, 0, 0), depth, cov2, rows=(240, 241), flags=flags)
    acc = flags[240] == 3
    assert acc.sum() > 500
    assert np.isnan(depth[240][acc]).all() and np.isnan(cov2[240][acc]).all()


def test_converges_to_ground_truth_distance():
    """Lateral sweep over the synthetic relief: the filter converges to the ray-cast |OP| distance
    (the quantity the maps hold, ref:299), like the RMS line of evaludateDepth ref:589."""
    seq = make_sequence("remode_640x480", n_frames=14)
    frames = [seq.render_host(i) for i in range(seq.n_frames)]
    _, gt = seq.render_host(0, with_distance=True)
    h, w = seq.shape
    depth = np.full((h, w), 3.0)
    cov2 = np.full((h, w), 3.0)
    rows = (200, 232)
    for i in range(1, seq.n_frames):
        T = seq.T_C_R(i)
        oracle.update(seq.params, frames[0], frames[i], T.q, T.t, depth, cov2, rows=rows, row_stride=4)
    ys = np.arange(rows[0], rows[1], 4)
    err = np.abs(depth[ys, 20:-20] - gt[ys, 20:-20])
    # 0.7 px sampling at a 5 cm baseline bounds the accuracy to a few cm at this point of the sweep
    assert np.median(err) < 0.03, f"median |depth - gt| = {np.median(err)}"
    assert np.median(cov2[ys, 20:-20]) < 0.05


def test_inverse_depth_variant_runs_and_differs():
    """USE_INVERSE_DEPTH_FOR_FILTERING ref:63,81-83,407-410,535-563."""
    seq = make_sequence("remode_640x480", n_frames=4, inverse_depth=True)
    frames = [seq.render_host(i) for i in range(seq.n_frames)]
    h, w = seq.shape
    depth = np.full((h, w), 3.0)
    cov2 = np.full((h, w), 0.5)  # init_cov2 of the inverse-depth arm, ref:272
    cnt = oracle.Counters()
    for i in range(1, 4):
        T = seq.T_C_R(i)
        oracle.update(seq.params, frames[0], frames[i], T.q, T.t, depth, cov2, rows=(240, 248), counters=cnt)
    assert cnt.accepted > 0.5 * cnt.active
    assert np.isfinite(depth[240:248, 20:-20]).mean() > 0.99


def test_pose_chain_matches_python_se3():
    """T_C_R = T_WC(curr)^-1 * T_WC(ref) ref:289-290: python host code == oracle, bit for bit."""
    rng = np.random.default_rng(5)
    L = oracle.lib()
    for _ in range(100):
        qa, qb = rng.normal(size=4), rng.normal(size=4)
        ta, tb = rng.normal(size=3), rng.normal(size=3)
        A = SE3.from_quat_trans(*qa, *ta)
        B = SE3.from_quat_trans(*qb, *tb)
        T = relative_pose(A, B)
        qo, to = (C.c_double * 4)(), (C.c_double * 3)()
        L.dmo_compose_T_C_R(_d(*qa), _d(*ta), _d(*qb), _d(*tb), qo, to)
        assert tuple(qo) == T.q and tuple(to) == T.t
        pt = rng.normal(size=3)
        out = (C.c_double * 3)()
        L.dmo_transform_point(_d(*T.q), _d(*T.t), _d(*pt), out)
        assert tuple(out) == T * pt


def test_evaluate_depth_and_mask():
    """evaludateDepth ref:569-590 and getMaskFromVariance ref:199-204."""
    rng = np.random.default_rng(6)
    p = oracle.default_params(640, 480)
    truth = rng.uniform(1, 3, (480, 640))
    est = truth + rng.normal(0, 0.01, (480, 640))
    var = 10.0 ** rng.uniform(-5, -3, (480, 640))
    s, n = C.c_double(), C.c_uint64()
    oracle.lib().dmo_evaluate_depth(C.byref(p), truth.ctypes.data, truth.strides[0], est.ctypes.data, est.strides[0],
                                    var.ctypes.data, var.strides[0], 2e-4, 0, 480, C.byref(s), C.byref(n))
    I = (slice(20, 460), slice(20, 620))
    m = var[I] < 2e-4
    assert n.value == m.sum()
    assert np.isclose(s.value, ((truth[I] - est[I])[m] ** 2).sum(), rtol=1e-12)
    mask = np.zeros((480, 640), np.uint8)
    var[5, 5] = np.nan
    oracle.lib().dmo_variance_mask(640, 480, var.ctypes.data, var.strides[0], 2e-4, mask.ctypes.data, 640)
    assert np.array_equal(mask, np.where(var > 2e-4, 0, 255).astype(np.uint8))
    assert mask[5, 5] == 255  # THRESH_BINARY_INV: NaN > thresh is false
