"""Oracle restatement vs the compiled reference translation unit, live (only where oracle/_ref was
built, i.e. in the container that has /root/reference; the frozen form of this check is
tests/test_oracle_golden.py)."""
import numpy as np
import pytest

import oracle


@pytest.fixture(scope="module")
def ref():
    L = oracle.ref_lib()
    if L is None:
        pytest.skip("oracle/_ref/libdmf_ref.so not built (no /root/reference here)")
    return L


def test_update_bit_equal_on_random_state(ref, seq640):
    """Random but plausible state maps (incl. converged, diverged and NaN pixels): one update() each."""
    seq, frames = seq640
    h, w = seq.shape
    rng = np.random.default_rng(7)
    for i in (1, 3, 5):
        depth = rng.uniform(1.5, 3.5, (h, w))
        cov2 = 10.0 ** rng.uniform(-5, 1.2, (h, w))  # spans min_cov=1e-4 .. max_cov=10
        depth[::37, ::41] = np.nan
        cov2[::53, ::29] = np.nan
        d2, c2 = depth.copy(), cov2.copy()
        T = seq.T_C_R(i)
        oracle.update(seq.params, frames[0], frames[i], T.q, T.t, depth, cov2)
        oracle.ref_update(frames[0], frames[i], T.q, T.t, d2, c2)
        assert np.array_equal(depth, d2, equal_nan=True)
        assert np.array_equal(cov2, c2, equal_nan=True)


def test_degenerate_poses_bit_equal(ref, seq640):
    """Zero baseline (acos(0/0) NaN poison, ref:527) and pure rotation."""
    seq, frames = seq640
    h, w = seq.shape
    for q, t in [((0, 0, 0, 1), (0, 0, 0)), ((0.01, -0.02, 0.005, 0.9997), (0, 0, 0)), ((0, 0, 0, 1), (1e-9, 0, 0))]:
        q = np.array(q, float)
        q /= np.linalg.norm(q)
        d1, c1 = np.full((h, w), 2.0), np.full((h, w), 0.5)
        d2, c2 = d1.copy(), c1.copy()
        oracle.update(seq.params, frames[0], frames[0], q, t, d1, c1, rows=(200, 216))
        oracle.ref_update(frames[0], frames[0], q, t, d2, c2)
        assert np.array_equal(d1[200:216], d2[200:216], equal_nan=True)
        assert np.array_equal(c1[200:216], c2[200:216], equal_nan=True)


# ---- "next" rows: the oracle legs pinned to the compiled reference -----------------------------------
def _five_updates(seq, frames):
    h, w = seq.shape
    d, c = np.full((h, w), 3.0), np.full((h, w), 3.0)
    for i in range(1, 6):
        T = seq.T_C_R(i)
        oracle.ref_update(frames[0], frames[i], T.q, T.t, d, c)
    return d, c


def test_evaluate_depth_mask_and_cloud_against_the_reference(ref, seq640):
    """evaludateDepth ref:569-590 (printed RMS captured at 17 digits), getMaskFromVariance ref:199-204, and
    getPointCloudFromImageAndDistance compiled from the reference's own utils/pointcloud/pointcloud_from_image_depth.h."""
    import ctypes as C
    seq, frames = seq640
    h, w = seq.shape
    _, gt = seq.render_host(0, with_distance=True)
    d, c = _five_updates(seq, frames)
    thr = float(np.median(c[20:-20, 20:-20]))
    po = oracle.to_params(seq.params)
    s, n = C.c_double(), C.c_uint64()
    oracle.lib().dmo_evaluate_depth(C.byref(po), gt.ctypes.data, gt.strides[0], d.ctypes.data, d.strides[0], c.ctypes.data,
                                    c.strides[0], thr, 0, h, C.byref(s), C.byref(n))
    rms_ref = oracle.ref_evaluate_depth(gt, d, c, thr)
    assert n.value > 1000 and np.isclose((s.value / n.value) ** 0.5, rms_ref, rtol=1e-12)  # OpenMP reduction order differs
    m_ref = oracle.ref_variance_mask(c, thr)
    m = np.zeros((h, w), np.uint8)
    oracle.lib().dmo_variance_mask(w, h, c.ctypes.data, c.strides[0], thr, m.ctypes.data, w)
    assert np.array_equal(m, m_ref) and 0 < (m == 255).mean() < 1
    color = np.ascontiguousarray(np.stack([frames[0], 255 - frames[0], frames[0] // 2], axis=-1))
    xyz_ref, rgb_ref = oracle.ref_point_cloud(color, d, m_ref)
    cap = (h - 40) * (w - 40)
    xyz, rgb = np.zeros((cap, 3), np.float32), np.zeros((cap, 3), np.uint8)
    k = oracle.lib().dmo_point_cloud(C.byref(po), color.ctypes.data, color.strides[0], 3, d.ctypes.data, d.strides[0],
                                     m.ctypes.data, w, xyz.ctypes.data, rgb.ctypes.data, cap)
    assert k == len(xyz_ref) > 1000
    assert np.array_equal(xyz[:k], xyz_ref) and np.array_equal(rgb[:k], rgb_ref)


def test_inverse_depth_arm_bit_equal_to_the_variant_translation_unit(seq640):
    """USE_INVERSE_DEPTH_FOR_FILTERING 1 (ref:63): oracle/Makefile generates the one-line-changed TU and compiles it."""
    from slamplay_b200.synth import make_sequence
    if oracle.ref_lib(inverse=True) is None:
        pytest.skip("oracle/_ref/libdmf_ref_inv.so not built (no /root/reference here)")
    _, frames = seq640
    seq = make_sequence("remode_640x480", n_frames=6, inverse_depth=True)
    h, w = seq.shape
    d1, c1 = np.full((h, w), 3.0), np.full((h, w), 0.5)  # init_cov2 = 0.5, ref:272
    d2, c2 = d1.copy(), c1.copy()
    for i in range(1, 6):
        T = seq.T_C_R(i)
        oracle.update(seq.params, frames[0], frames[i], T.q, T.t, d1, c1)
        oracle.ref_update(frames[0], frames[i], T.q, T.t, d2, c2, inverse=True)
        assert np.array_equal(d1, d2, equal_nan=True) and np.array_equal(c1, c2, equal_nan=True), i
    assert (c1 != 0.5).mean() > 0.5


def test_dataset_reader_and_pose_chain_against_the_reference(ref, seq640, tmp_path):
    """readDatasetFiles ref:317-352 on a REMODE-layout directory written by slamplay_b200.remode.write_dataset, and the
    pose chain T_C_R = T_WC(i)^-1 * T_WC(0) ref:289-290 — Python reader / se3 and the oracle against the compiled TU."""
    import ctypes as C

    from slamplay_b200.remode import read_dataset, write_dataset
    from slamplay_b200.se3 import relative_pose
    seq, frames = seq640
    _, gt = seq.render_host(0, with_distance=True)
    write_dataset(str(tmp_path), seq, frames, gt)
    files_ref, poses_ref, depth_ref = oracle.ref_read_dataset(str(tmp_path))
    files, poses, depth = read_dataset(str(tmp_path))
    n = len(files)
    # the reference's `while (!fin.eof())` loop appends one bogus entry after a trailing newline (its driver skips it:
    # imread fails, ref:288); the complete entries must agree exactly
    assert n == seq.n_frames and len(files_ref) in (n, n + 1)
    assert files_ref[:n] == files
    for k in range(n):
        assert tuple(poses_ref[k][:4]) == tuple(poses[k].q) and tuple(poses_ref[k][4:]) == tuple(poses[k].t)
    assert np.array_equal(depth, depth_ref)
    for k in range(1, n):
        T_ref = oracle.ref_compose_T_C_R(poses_ref[0], poses_ref[k])
        T = relative_pose(poses[0], poses[k])
        assert tuple(T_ref[:4]) == tuple(T.q) and tuple(T_ref[4:]) == tuple(T.t)
        qo, to = (C.c_double * 4)(), (C.c_double * 3)()
        oracle.lib().dmo_compose_T_C_R((C.c_double * 4)(*poses[0].q), (C.c_double * 3)(*poses[0].t), (C.c_double * 4)(*poses[k].q),
                                       (C.c_double * 3)(*poses[k].t), qo, to)
        assert tuple(T_ref[:4]) == tuple(qo) and tuple(T_ref[4:]) == tuple(to)
