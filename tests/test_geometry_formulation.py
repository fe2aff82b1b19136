"""The FP64 geometry of the CUDA path (slamplay_b200/csrc/dmf_geometry.h: epipolar segment, accumulated sample
positions, trip count, 2x2 ColPivHouseholderQR triangulation, Gaussian fusion) compiled for the HOST and compared
BIT FOR BIT with the oracle, which is itself pinned bit for bit to the compiled reference translation unit
(tests/test_oracle_vs_ref.py).  On the GPU the same source runs with __dmul_rn / __dadd_rn / __ddiv_rn / __dsqrt_rn
(never contracted, IEEE-rounded); only acos / sin (CUDA libm vs glibc) can differ there, by an ulp.

Includes the case that motivated the rewrite: pixels next to the epipole of a frame, where the triangulation system of
ref:505-516 is singular and a Cramer solve returns 1e4 where Eigen's column-pivoted QR returns 1e-3."""
import ctypes as C
import subprocess
from pathlib import Path

import numpy as np
import pytest

import oracle
from slamplay_b200.synth import make_params, make_sequence

HERE = Path(__file__).resolve().parent
G = HERE / "golden"


@pytest.fixture(scope="module")
def gh():
    src = HERE / "geom_host" / "geom_host.cpp"
    out = HERE / "geom_host" / "libgeom_host.so"
    hdr = HERE.parent / "slamplay_b200" / "csrc" / "dmf_geometry.h"
    if not out.exists() or out.stat().st_mtime < max(src.stat().st_mtime, hdr.stat().st_mtime):
        subprocess.run(["/usr/bin/g++" if Path("/usr/bin/g++").exists() else "g++", "-O2", "-march=x86-64-v3", "-ffp-contract=off",
                        "-fPIC", "-shared", "-std=c++17", "-o", str(out), str(src)], check=True)
    L = C.CDLL(str(out))
    d = C.c_double
    P = C.POINTER(d)
    L.gh_search.argtypes = [P, P, P, d, d, d, d, d, d, d, d, C.c_int, P]
    L.gh_sample.argtypes = [P, P, d, d, C.c_int, P]
    L.gh_trip_count.argtypes = [d, d]
    L.gh_trip_count.restype = C.c_int
    L.gh_fuse.argtypes = [P, P, P, d, d, d, d, d, d, d, d, d, C.c_int, P]
    L.gh_qr_solve2.argtypes = [P, P, P]
    return L


def arr(*v):
    return (C.c_double * len(v))(*[float(x) for x in v])


def host_inverse(q, t):
    """dmf_api.cu se3_inverse (Sophus SE3::inverse): conjugate, normalised; t' = R^-1 * (t * -1); |t'| in Eigen's order."""
    qi, ti = (C.c_double * 4)(), (C.c_double * 3)()
    oracle_inverse = np.zeros(7)
    c = np.array([-q[0], -q[1], -q[2], q[3]])
    n = np.sqrt((c[0] * c[0] + c[2] * c[2]) + (c[1] * c[1] + c[3] * c[3]))
    c = c / n
    v = np.array([t[0] * -1.0, t[1] * -1.0, t[2] * -1.0])
    u = np.array([c[1] * v[2] - c[2] * v[1], c[2] * v[0] - c[0] * v[2], c[0] * v[1] - c[1] * v[0]])
    u = u + u
    cr = np.array([c[1] * u[2] - c[2] * u[1], c[2] * u[0] - c[0] * u[2], c[0] * u[1] - c[1] * u[0]])
    ti_ = (v + c[3] * u) + cr
    tn = np.sqrt(ti_[0] * ti_[0] + (ti_[1] * ti_[1] + ti_[2] * ti_[2]))
    return arr(*c), arr(*ti_), float(tn)


def check_search_and_fuse(gh, p, q, t, cases, inverse=False):
    """cases: (x, y, mu, sigma2).  Every number the kernels derive must equal the oracle's bits."""
    L = oracle.lib()
    po = oracle.to_params(p)
    cam = arr(p.fx, p.fy, p.cx, p.cy)
    qa, ta = arr(*q), arr(*t)
    qi, ti, tn = host_inverse(q, t)
    img = np.zeros((p.height, p.width), np.uint8)  # flat images: every sample ties at NCC 0, the first one wins
    n_checked = 0
    for (x, y, mu, s2) in cases:
        sigma = float(np.sqrt(s2))
        want = (C.c_double * 9)()
        L.dmo_epipolar_search(C.byref(po), img.ctypes.data, img.strides[0], img.ctypes.data, img.strides[0], qa, ta, x, y, mu, sigma, want)
        got = (C.c_double * 6)()
        gh.gh_search(cam, qa, ta, x, y, mu, sigma, p.n_sigma, p.min_depth, p.max_half_len, p.step, int(inverse), got)
        assert (got[2], got[3]) == (want[3], want[4]) or (np.isnan(got[2]) and np.isnan(want[3])), (x, y, "direction")
        assert int(got[5]) == int(want[7]), (x, y, "trip count", got[5], want[7], got[4])
        # the oracle reports the winning sample's position: with flat images that is the first sample inside the border
        if want[6] > 0:
            k = int(want[8])
            pos = (C.c_double * 2)()
            gh.gh_sample(arr(got[0], got[1]), arr(got[2], got[3]), got[4], p.step, k, pos)
            assert (pos[0], pos[1]) == (want[1], want[2]), (x, y, "sample position", k)
        # fusion at a few sample positions of this segment, including the far end (behind-camera solutions)
        n = int(got[5])
        for k in sorted({0, n // 3, n // 2, max(n - 1, 0)}):
            if n == 0:
                continue
            pos = (C.c_double * 2)()
            gh.gh_sample(arr(got[0], got[1]), arr(got[2], got[3]), got[4], p.step, k, pos)
            fo = (C.c_double * 4)()
            L.dmo_update_depth_filter(C.byref(po), qa, ta, x, y, pos[0], pos[1], got[2], got[3], mu, s2, fo)
            fg = (C.c_double * 4)()
            gh.gh_fuse(cam, qi, ti, tn, x, y, pos[0], pos[1], got[2], got[3], mu, s2, int(inverse), fg)
            for a, b, name in zip(fg, fo, ("depth_est", "d_cov2", "mu_fuse", "sigma_fuse2")):
                assert a == b or (np.isnan(a) and np.isnan(b)), (x, y, k, name, a, b)
            n_checked += 1
    return n_checked


def test_golden_unit_cases_bit_equal(gh):
    u = np.load(G / "remode640_ref_units.npz")
    p = make_params(640, 480)
    cases = [(u["rx"][i], u["ry"][i], u["mu"][i], u["sigma"][i] ** 2) for i in range(len(u["rx"]))]
    assert check_search_and_fuse(gh, p, u["pose"][:4], u["pose"][4:], cases) > 500


@pytest.mark.parametrize("workload,frame", [("uhd_3840x2160", 1), ("hd_1920x1080", 8), ("kitti_1241x376", 5), ("remode_640x480", 40)])
def test_sequence_poses_including_the_epipole_neighbourhood(gh, workload, frame):
    """Random pixels and states under the poses of the benchmark sequences, plus a dense patch around the epipole of the
    frame (rays parallel to the baseline: the 2x2 system is singular; at 4K frame 1 a Cramer solve gives depth 1e4 where
    the reference's QR gives 1e-3)."""
    seq = make_sequence(workload, n_frames=frame + 1)
    p = seq.params
    T = seq.T_C_R(frame)
    rng = np.random.default_rng(5)
    b = p.border
    cases = [(float(rng.integers(b, p.width - b)), float(rng.integers(b, p.height - b)), float(rng.uniform(0.5, 4.0)),
              float(10.0 ** rng.uniform(-4, 0.9))) for _ in range(300)]
    # epipole of the frame in the reference image: projection of the current camera centre t_RC
    qi, ti, _ = host_inverse(T.q, T.t)
    if abs(ti[2]) > 1e-12:
        ex, ey = p.fx * ti[0] / ti[2] + p.cx, p.fy * ti[1] / ti[2] + p.cy
        for dx in range(-6, 7, 2):
            for dy in range(-6, 7, 2):
                x, y = round(ex) + dx, round(ey) + dy
                if b <= x < p.width - b and b <= y < p.height - b:
                    cases.append((float(x), float(y), 3.0, 3.0))
                    cases.append((float(x), float(y), float(rng.uniform(1.0, 3.0)), float(10.0 ** rng.uniform(-3, 0))))
    assert check_search_and_fuse(gh, p, T.q, T.t, cases) > 500


def test_inverse_depth_arm_bit_equal(gh):
    seq = make_sequence("remode_640x480", n_frames=6, inverse_depth=True)
    rng = np.random.default_rng(9)
    cases = [(float(rng.integers(20, 620)), float(rng.integers(20, 460)), float(rng.uniform(1.0, 4.0)), float(10.0 ** rng.uniform(-4, -0.1)))
             for _ in range(200)]
    assert check_search_and_fuse(gh, seq.params, seq.T_C_R(4).q, seq.T_C_R(4).t, cases, inverse=True) > 300


def test_degenerate_poses_bit_equal(gh):
    """Zero baseline (acos(0/0) NaN poison, ref:527), pure rotation, NaN state: same NaNs, same zero-length segments."""
    p = make_params(640, 480)
    cases = [(100.0, 100.0, 2.0, 0.5), (320.0, 240.0, 3.0, 3.0), (500.0, 400.0, float("nan"), 1.0), (50.0, 450.0, 1.0, float("nan"))]
    for q, t in [((0, 0, 0, 1), (0, 0, 0)), ((0.01, -0.02, 0.005, 0.9997), (0, 0, 0)), ((0, 0, 0, 1), (1e-9, 0, 0))]:
        q = np.array(q, float)
        q /= np.linalg.norm(q)
        check_search_and_fuse(gh, p, q, t, cases)


def test_trip_count_matches_the_accumulated_loop(gh):
    """`for (l = -half; l <= half; l += 0.7)` ref:432: closed form away from boundaries, the loop next to them."""
    rng = np.random.default_rng(3)
    halves = list(rng.uniform(0, 100, 2000)) + [0.0, 100.0, 0.35, 0.7, 1.05, 3.5, 7.0, 35.0, 70.0, 0.35 * 57, np.nextafter(0.35, 1), np.nextafter(0.35, 0)]
    halves += [0.35 * k for k in range(1, 286)] + [float(np.nextafter(0.35 * k, 0)) for k in range(1, 286)] + [float(np.nextafter(0.35 * k, 1e9)) for k in range(1, 286)]
    for h in halves:
        n, l = 0, -h
        while l <= h:
            n += 1
            l += 0.7
        assert gh.gh_trip_count(h, 0.7) == n, h
    assert gh.gh_trip_count(float("nan"), 0.7) == 0


def test_qr_solve_matches_the_oracle_restatement(gh):
    rng = np.random.default_rng(17)
    L = oracle.lib()
    for i in range(3000):
        c = 1.0 - 10.0 ** rng.uniform(-17, 0)  # f_ref . f2 from orthogonal to parallel
        a = np.array([1.0 + rng.normal() * 1e-16, -c, c, -(1.0 + rng.normal() * 1e-16)])
        if i % 7 == 0:
            a = rng.normal(size=4)
        if i % 97 == 0:
            a[:] = 0
        b = rng.normal(size=2) * 10.0 ** rng.uniform(-6, 0)
        want, got = (C.c_double * 2)(), (C.c_double * 2)()
        L.dmo_qr_solve2(arr(*a), arr(*b), want)
        gh.gh_qr_solve2(arr(*a), arr(*b), got)
        assert (got[0], got[1]) == (want[0], want[1]) or (np.isnan(got[0]) and np.isnan(want[0])), (a, b)
