"""Parity of the CUDA path (through the C ABI) against the CPU oracle — the first gate.

Tolerances (BASELINE.json north_star): depth within 1e-3 relative on >= 95 % of interior pixels;
accept/reject (NCC >= 0.85f, ref:443) and covariance-gate (ref:366) decisions differing on <= 0.5 % of
pixels.  Measured margins are far tighter (1e-6 / 0 mismatches); tighter secondary asserts keep it so."""
import ctypes as C
import hashlib
from pathlib import Path

import numpy as np
import pytest

import oracle
from parity import (MAX_DECISION_MISMATCH, MIN_DEPTH_AGREE, class_mismatch, depth_agreement, flag_mismatch)
from slamplay_b200.se3 import SE3
from slamplay_b200.synth import make_sequence

pytestmark = pytest.mark.gpu
G = Path(__file__).resolve().parent / "golden"


@pytest.fixture(scope="module")
def DF():
    from slamplay_b200.depth_filter import DepthFilter
    return DepthFilter


def run_pair(DF, seq, frames, n_updates, rows=None, stride=1, init=(3.0, 3.0), state=None, check_flags=True):
    """GPU and oracle side by side; returns final maps of both and the worst per-frame decision mismatch."""
    p = seq.params
    h, w = seq.shape
    f = DF(p)
    f.set_reference(frames[0])
    if state is None:
        f.fill_state(*init)
        d_ref, c_ref = np.full((h, w), init[0]), np.full((h, w), init[1])
    else:
        d_ref, c_ref = state[0].copy(), state[1].copy()
        f.upload_state(d_ref, c_ref)
    f.enable_flags(check_flags)
    r0, r1 = rows if rows else (p.border, h - p.border)
    ys = range(r0, r1, stride)
    worst = 0.0
    oc = oracle.Counters()
    for i in range(1, n_updates + 1):
        T = seq.T_C_R(i)
        f.update(frames[i], T)
        fl_ref = np.zeros((h, w), np.uint8)
        oracle.update(p, frames[0], frames[i], T.q, T.t, d_ref, c_ref, rows=(r0, r1), row_stride=stride, flags=fl_ref, counters=oc)
        if check_flags:
            worst = max(worst, flag_mismatch(p, f.flags(), fl_ref, ys))
    d, c = f.download_state()
    cnt = f.counters()
    f.close()
    return d, c, d_ref, c_ref, worst, ys, cnt, oc.as_dict()


def test_remode640_sequence(DF, seq640):
    """BASELINE.json config 2 (shortened): 640x480 REMODE-shaped, every interior pixel, 5 updates."""
    seq, frames = seq640
    d, c, d_ref, c_ref, worst, ys, cnt, oc = run_pair(DF, seq, frames, 5)
    p = seq.params
    agree = depth_agreement(p, d, d_ref)
    assert agree >= MIN_DEPTH_AGREE, agree
    assert worst <= MAX_DECISION_MISMATCH and class_mismatch(p, c, c_ref) <= MAX_DECISION_MISMATCH
    # observed margins: keep them
    assert depth_agreement(p, d, d_ref, rtol=1e-6) > 0.9999
    assert worst < 1e-4
    for k in ("interior", "active", "ncc_evals", "accepted"):
        assert abs(cnt[k] - oc[k]) <= 1e-4 * max(oc[k], 1), (k, cnt[k], oc[k])
    # border pixels are never touched (ref:357,363)
    assert (d[:p.border] == 3.0).all() and (d[:, :p.border] == 3.0).all() and (c[-p.border:] == 3.0).all()


@pytest.mark.parametrize("w,h,border", [(64, 64, 20), (97, 75, 13), (130, 67, 13), (333, 201, 20)])
def test_ragged_sizes_and_minimum_border(DF, w, h, border):
    """Smallest frame, odd widths / heights (partial 32x8 tiles, generic-width ncc_kernel, per-column moments kernel) and
    the smallest border the tables support (13: samples read block positions up to x = width-16, y = height-9)."""
    seq = make_sequence("custom", n_frames=5, width=w, height=h, kind="lateral")
    seq.params.border = border
    frames = [seq.render_host(i) for i in range(seq.n_frames)]
    d, c, d_ref, c_ref, worst, ys, cnt, oc = run_pair(DF, seq, frames, 4)
    p = seq.params
    assert oc["ncc_evals"] > 0 and oc["interior"] == 4 * (w - 2 * border) * (h - 2 * border)
    assert depth_agreement(p, d, d_ref, rtol=1e-6) > 0.999
    assert worst <= MAX_DECISION_MISMATCH and class_mismatch(p, c, c_ref) <= MAX_DECISION_MISMATCH
    for k in ("interior", "active", "ncc_evals", "accepted"):
        assert abs(cnt[k] - oc[k]) <= 1e-3 * max(oc[k], 1), (k, cnt[k], oc[k])
    assert (d[:border] == 3.0).all() and (d[:, :border] == 3.0).all() and (c[-border:] == 3.0).all() and (c[:, -border:] == 3.0).all()


def test_golden_sequence_from_compiled_reference(DF):
    """GPU vs the fixture produced by the compiled reference TU (tests/golden/make_golden.py)."""
    g = np.load(G / "remode640_ref_update.npz")
    n = int(g["n_frames"])
    seq = make_sequence("remode_640x480", n_frames=n)
    frames = [seq.render_host(i) for i in range(n)]
    assert [hashlib.sha256(f.tobytes()).hexdigest() for f in frames] == list(g["frame_sha"])
    f = DF(seq.params)
    f.set_reference(frames[0])
    f.fill_state(3.0, 3.0)
    for i in range(1, n):
        f.update(frames[i], (g["poses"][i - 1][:4], g["poses"][i - 1][4:]))
    d, c = f.download_state()
    f.close()
    step = int(g["row_step"])
    rows = range(0, 480, step)
    p = seq.params
    assert depth_agreement(p, d, _expand(g["depth_rows"], step, d.shape), rows) >= MIN_DEPTH_AGREE
    assert depth_agreement(p, d, _expand(g["depth_rows"], step, d.shape), rows, rtol=1e-6) > 0.9999
    assert class_mismatch(p, c, _expand(g["cov2_rows"], step, c.shape), rows) <= MAX_DECISION_MISMATCH


def _expand(rows_arr, step, shape):
    full = np.zeros(shape)
    full[::step] = rows_arr
    return full


def test_ncc_values_against_oracle(DF, seq640):
    """Best NCC per pixel (ref:430-441) from the integer-moment kernel vs the FP64 two-pass oracle."""
    seq, frames = seq640
    p = seq.params
    h, w = seq.shape
    f = DF(p)
    f.set_reference(frames[0])
    f.fill_state(3.0, 3.0)
    f.enable_flags(True)
    T = seq.T_C_R(2)
    f.update(frames[2], T)
    ncc, trips, k = f.debug()
    d, c = np.full((h, w), 3.0), np.full((h, w), 3.0)
    oncc = np.zeros((h, w), np.float32)
    on = np.zeros((h, w), np.int32)
    oracle.update(p, frames[0], frames[2], T.q, T.t, d, c, dbg_ncc=oncc, dbg_n=on)
    I = (slice(20, h - 20), slice(20, w - 20))
    assert np.abs(ncc[I] - oncc[I]).max() < 2e-6
    assert (trips[I] >= on[I]).all()  # trip count >= NCC calls (samples outside the border are skipped)
    f.close()


def test_kitti_shaped_forward_motion(DF):
    """BASELINE.json config 3 (shortened): 1241x376 (odd width -> padded pitch), radial epipolar lines."""
    seq = make_sequence("kitti_1241x376", n_frames=5)
    frames = [seq.render_host(i) for i in range(seq.n_frames)]
    d, c, d_ref, c_ref, worst, ys, cnt, oc = run_pair(DF, seq, frames, 4, stride=6)
    p = seq.params
    assert depth_agreement(p, d, d_ref, ys) >= MIN_DEPTH_AGREE
    assert depth_agreement(p, d, d_ref, ys, rtol=1e-6) > 0.999
    assert worst <= MAX_DECISION_MISMATCH and class_mismatch(p, c, c_ref, ys) <= MAX_DECISION_MISMATCH


def test_inverse_depth_variant(DF):
    """USE_INVERSE_DEPTH_FOR_FILTERING (ref:63): thresholds ref:82-83, init cov 0.5 ref:272."""
    seq = make_sequence("remode_640x480", n_frames=4, inverse_depth=True)
    frames = [seq.render_host(i) for i in range(seq.n_frames)]
    d, c, d_ref, c_ref, worst, ys, _, _ = run_pair(DF, seq, frames, 3, stride=4, init=(3.0, 0.5))
    p = seq.params
    assert depth_agreement(p, d, d_ref, ys) >= MIN_DEPTH_AGREE
    assert worst <= MAX_DECISION_MISMATCH and class_mismatch(p, c, c_ref, ys) <= MAX_DECISION_MISMATCH


def test_edge_states_nan_converged_diverged(DF, seq640):
    """ref:366: NaN passes the gate and stays NaN; converged / diverged pixels are bit-identical afterwards."""
    seq, frames = seq640
    h, w = seq.shape
    rng = np.random.default_rng(11)
    depth = rng.uniform(1.5, 3.5, (h, w))
    cov2 = 10.0 ** rng.uniform(-5, 1.2, (h, w))
    depth[::37, ::41] = np.nan
    cov2[::53, ::29] = np.nan
    d, c, d_ref, c_ref, worst, ys, _, _ = run_pair(DF, seq, frames, 2, state=(depth, cov2))
    p = seq.params
    assert worst <= MAX_DECISION_MISMATCH
    assert depth_agreement(p, d, d_ref) >= MIN_DEPTH_AGREE
    skip = (cov2 < p.min_cov) | (cov2 > p.max_cov)
    assert np.array_equal(d[skip], depth[skip], equal_nan=True) and np.array_equal(c[skip], cov2[skip], equal_nan=True)
    assert np.isnan(c[::53, ::29]).all()
    nan_d = np.isnan(depth) & ~skip
    assert np.isnan(d[nan_d][:]).all()  # NaN depth -> zero samples -> stays NaN


def test_zero_baseline_nan_poison(DF, seq640):
    """|t| = 0: acos(0/0) (ref:527) poisons every accepted pixel with NaN, in the reference and here."""
    seq, frames = seq640
    p = seq.params
    h, w = seq.shape
    f = DF(p)
    f.set_reference(frames[0])
    f.fill_state(2.0, 0.5)
    f.enable_flags(True)
    f.update(frames[0], SE3.identity())
    fl = f.flags()
    d, c = f.download_state()
    f.close()
    d_ref, c_ref = np.full((h, w), 2.0), np.full((h, w), 0.5)
    fl_ref = np.zeros((h, w), np.uint8)
    oracle.update(p, frames[0], frames[0], (0, 0, 0, 1), (0, 0, 0), d_ref, c_ref, flags=fl_ref)
    assert flag_mismatch(p, fl, fl_ref) <= MAX_DECISION_MISMATCH
    I = (slice(20, h - 20), slice(20, w - 20))
    assert (np.isnan(d[I]) == np.isnan(d_ref[I])).mean() > 0.995
    assert np.isnan(d[I]).mean() > 0.9


def test_strict_dropin_update_function(DF, seq640):
    """The reference's free function: update(ref, curr, T_C_R, depth, depth_cov2) in place (ref:355)."""
    from slamplay_b200.depth_filter import release_strict_contexts, update
    seq, frames = seq640
    h, w = seq.shape
    depth, cov2 = np.full((h, w), 3.0), np.full((h, w), 3.0)
    d_ref, c_ref = depth.copy(), cov2.copy()
    for i in (1, 2):
        T = seq.T_C_R(i)
        update(frames[0], frames[i], T, depth, cov2)
        oracle.update(seq.params, frames[0], frames[i], T.q, T.t, d_ref, c_ref)
    release_strict_contexts()
    assert depth_agreement(seq.params, depth, d_ref, rtol=1e-6) > 0.9999
    with pytest.raises(ValueError):
        update(frames[0][:100], frames[1], seq.T_C_R(1), depth, cov2)  # MSG_ASSERT-like size check (ref:265)


def test_strided_host_buffers_and_unaligned_device_frames(DF, seq640):
    """cv::Mat::step larger than the row (ROI views) and a device frame whose pointer is not 4-byte aligned."""
    import torch
    seq, frames = seq640
    p = seq.params
    h, w = seq.shape
    big = np.zeros((h, w + 37), np.uint8)
    big[:, 5:5 + w] = frames[2]
    view = big[:, 5:5 + w]
    dbig = np.full((h, w + 3), 3.0)
    cbig = np.full((h, w + 3), 3.0)
    f = DF(p)
    f.set_reference(frames[0])
    f.upload_state(dbig[:, 1:1 + w], cbig[:, 1:1 + w])
    T = seq.T_C_R(2)
    f.update(view, T)
    d1, c1 = f.download_state()
    # same update from an odd device address
    f.fill_state(3.0, 3.0)
    dev = torch.zeros(h * w + 16, dtype=torch.uint8, device="cuda")
    dev[1:1 + h * w] = torch.from_numpy(frames[2].reshape(-1)).cuda()
    torch.cuda.synchronize()
    f.update_device(dev.data_ptr() + 1, w, T)
    d2, c2 = f.download_state()
    f.close()
    assert np.array_equal(d1, d2, equal_nan=True) and np.array_equal(c1, c2, equal_nan=True)
    d_ref, c_ref = np.full((h, w), 3.0), np.full((h, w), 3.0)
    oracle.update(p, frames[0], frames[2], T.q, T.t, d_ref, c_ref, rows=(20, 460), row_stride=8)
    assert depth_agreement(p, d1, d_ref, range(20, 460, 8), rtol=1e-6) > 0.9999


def test_band_contexts_are_bit_identical_to_one_context(DF, seq640):
    """Row-band sharding invariant: results do not depend on the partition."""
    seq, frames = seq640
    p = seq.params
    h, w = seq.shape
    whole = DF(p)
    bands = [DF(p, rows=(0, 200)), DF(p, rows=(200, 333)), DF(p, rows=(333, h))]
    for f in [whole] + bands:
        f.set_reference(frames[0])
        f.fill_state(3.0, 3.0)
    for i in (1, 2, 3):
        for f in [whole] + bands:
            f.update(frames[i], seq.T_C_R(i))
    d, c = whole.download_state()
    d2, c2 = np.zeros((h, w)), np.zeros((h, w))
    tot = {"active": 0, "ncc_evals": 0, "accepted": 0, "interior": 0}
    for f in bands:
        f.download_state(d2, c2)
        for k in tot:
            tot[k] += f.counters()[k]
    cw = whole.counters()
    for f in [whole] + bands:
        f.close()
    assert np.array_equal(d, d2, equal_nan=True) and np.array_equal(c, c2, equal_nan=True)
    assert all(tot[k] == cw[k] for k in tot)


@pytest.mark.parametrize("blk", [16, 8])  # 28 blocks: the incomplete last round is odd; 55 blocks: even (dealt in reverse too)
def test_cyclic_contexts_are_bit_identical_to_one_context(DF, seq640, blk):
    """Block-cyclic row ownership (dmf_create_cyclic): three interleaved contexts == one context."""
    from slamplay_b200.sharded import cyclic_rows
    seq, frames = seq640
    p = seq.params
    h, w = seq.shape
    whole = DF(p)
    parts = [DF(p, cyclic=(blk, 3, r)) for r in range(3)]
    rows = np.concatenate([f.owned_rows() for f in parts])
    assert sorted(rows.tolist()) == list(range(p.border, h - p.border))
    for r, f in enumerate(parts):  # the host-side replica of the dealing rule (gather / scatter of the sharded filter)
        assert np.array_equal(np.sort(f.owned_rows()), cyclic_rows(h, p.border, blk, 3, r))
    for f in [whole] + parts:
        f.set_reference(frames[0])
        f.fill_state(3.0, 3.0)
    for i in (1, 2, 3):
        for f in [whole] + parts:
            f.update(frames[i], seq.T_C_R(i))
    d, c = whole.download_state()
    d2, c2 = np.full((h, w), 3.0), np.full((h, w), 3.0)
    tot = {"active": 0, "ncc_evals": 0, "accepted": 0, "interior": 0}
    for f in parts:
        f.download_state(d2, c2)
        for k in tot:
            tot[k] += f.counters()[k]
    cw = whole.counters()
    for f in [whole] + parts:
        f.close()
    assert np.array_equal(d, d2, equal_nan=True) and np.array_equal(c, c2, equal_nan=True)
    assert all(tot[k] == cw[k] for k in tot)


def test_full_size_properties_hd(DF):
    """BASELINE.json config 4 size (1920x1080): determinism (two runs bit-identical), exact parity on a row
    subset against the oracle, work counters consistent."""
    import torch
    seq = make_sequence("hd_1920x1080", n_frames=5)
    p = seq.params
    h, w = seq.shape
    pitch = (w + 15) // 16 * 16
    dev = torch.zeros((seq.n_frames, h, pitch), dtype=torch.uint8, device="cuda")
    for i in range(seq.n_frames):
        seq.render_device(i, dev[i].data_ptr(), pitch, stream=torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    outs = []
    for _ in range(2):
        f = DF(p)
        f.set_reference_device(dev[0].data_ptr(), pitch)
        f.fill_state(3.0, 3.0)
        for i in range(1, seq.n_frames):
            f.update_device(dev[i].data_ptr(), pitch, seq.T_C_R(i))
        outs.append(f.download_state() + (f.counters(),))
        f.close()
    assert np.array_equal(outs[0][0], outs[1][0], equal_nan=True) and np.array_equal(outs[0][1], outs[1][1], equal_nan=True)
    cnt = outs[0][2]
    assert cnt["interior"] == 4 * (h - 40) * (w - 40) and cnt["accepted"] <= cnt["active"] <= cnt["interior"]
    host = dev[:, :, :w].cpu().numpy()
    rows = (20, h - 20)
    d_ref, c_ref = np.full((h, w), 3.0), np.full((h, w), 3.0)
    for i in range(1, seq.n_frames):
        T = seq.T_C_R(i)
        oracle.update(p, host[0], host[i], T.q, T.t, d_ref, c_ref, rows=rows, row_stride=64)
    ys = range(20, h - 20, 64)
    assert depth_agreement(p, outs[0][0], d_ref, ys) >= MIN_DEPTH_AGREE
    assert depth_agreement(p, outs[0][0], d_ref, ys, rtol=1e-6) > 0.99
    assert class_mismatch(p, outs[0][1], c_ref, ys) <= MAX_DECISION_MISMATCH


def test_gpu_renderer_is_bit_identical_to_cpu_renderer():
    import torch
    seq = make_sequence("tiny", width=320, height=240, n_frames=3)
    h, w = seq.shape
    img = torch.zeros((h, w), dtype=torch.uint8, device="cuda")
    dist = torch.zeros((h, w), dtype=torch.float64, device="cuda")
    seq.render_device(2, img.data_ptr(), w, dist.data_ptr(), w * 8, stream=torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    cpu_img, cpu_dist = seq.render_host(2, with_distance=True)
    assert np.array_equal(img.cpu().numpy(), cpu_img)
    assert np.array_equal(dist.cpu().numpy(), cpu_dist)


def test_evaluate_depth_and_variance_mask(DF, seq640):
    """'next' rows SURVEY.md §8f: evaludateDepth ref:569-590 and getMaskFromVariance ref:199-204 on the device."""
    seq, frames = seq640
    p = seq.params
    h, w = seq.shape
    _, gt = seq.render_host(0, with_distance=True)
    f = DF(p)
    f.set_reference(frames[0])
    f.fill_state(3.0, 3.0)
    for i in range(1, 6):
        f.update(frames[i], seq.T_C_R(i))
    f.set_truth(gt)
    d, c = f.download_state()
    thr = float(np.median(c[20:-20, 20:-20]))
    s, n = f.evaluate_depth(thr)
    mask = f.variance_mask(thr)
    f.close()
    so, no = C.c_double(), C.c_uint64()
    po = oracle.to_params(p)
    oracle.lib().dmo_evaluate_depth(C.byref(po), gt.ctypes.data, gt.strides[0], d.ctypes.data, d.strides[0], c.ctypes.data,
                                    c.strides[0], thr, 0, h, C.byref(so), C.byref(no))
    assert n == no.value and np.isclose(s, so.value, rtol=1e-10)
    m_ref = np.zeros((h, w), np.uint8)
    oracle.lib().dmo_variance_mask(w, h, c.ctypes.data, c.strides[0], thr, m_ref.ctypes.data, w)
    assert np.array_equal(mask, m_ref)


def test_point_cloud_export(DF, seq640):
    """'next' row: getPointCloudFromImageAndDistance (utils/pointcloud/pointcloud_from_image_depth.h:42-89)."""
    seq, frames = seq640
    p = seq.params
    h, w = seq.shape
    f = DF(p)
    f.set_reference(frames[0])
    f.fill_state(3.0, 3.0)
    for i in range(1, 6):
        f.update(frames[i], seq.T_C_R(i))
    d, c = f.download_state()
    thr = float(np.median(c[20:-20, 20:-20]))
    color = np.stack([frames[0], 255 - frames[0], frames[0] // 2], axis=-1)  # BGR
    xyz, rgb = f.point_cloud(color, thr)
    f.close()
    mask = np.zeros((h, w), np.uint8)
    oracle.lib().dmo_variance_mask(w, h, c.ctypes.data, c.strides[0], thr, mask.ctypes.data, w)
    cap = (h - 40) * (w - 40)
    xo = np.zeros((cap, 3), np.float32)
    ro = np.zeros((cap, 3), np.uint8)
    po = oracle.to_params(p)
    n = oracle.lib().dmo_point_cloud(C.byref(po), color.ctypes.data, color.strides[0], 3, d.ctypes.data, d.strides[0],
                                     mask.ctypes.data, w, xo.ctypes.data, ro.ctypes.data, cap)
    assert n == len(xyz) and 0.3 * cap < n < 0.7 * cap
    assert np.array_equal(rgb, ro[:n])
    assert np.allclose(xyz, xo[:n], rtol=1e-6, atol=0)  # float32 of the same FP64 values (ulp-level differences)


def test_api_errors(DF, seq640):
    from slamplay_b200.depth_filter import DmfError
    seq, frames = seq640
    f = DF(seq.params)
    with pytest.raises(DmfError, match="set_reference"):
        f.update(frames[1], seq.T_C_R(1))
    with pytest.raises(ValueError):
        f.set_reference(frames[0].astype(np.float32))
    with pytest.raises(ValueError, match="max_cov"):
        f.fill_state(3.0, 11.0)  # MSG_ASSERT(init_cov2 < max_cov) ref:276
    f.close()


def test_sharded_filter_world1(DF, seq640):
    """ShardedDepthFilter with a single rank is the plain resident filter."""
    import torch
    from slamplay_b200.sharded import ShardedDepthFilter
    seq, frames = seq640
    h, w = seq.shape
    pitch = (w + 15) // 16 * 16
    dev = torch.zeros((4, h, pitch), dtype=torch.uint8, device="cuda")
    for i in range(4):
        dev[i, :, :w] = torch.from_numpy(frames[i]).cuda()
    torch.cuda.synchronize()
    sf = ShardedDepthFilter(seq.params, device=0)
    sf.set_reference(dev[0])
    sf.fill_state(3.0, 3.0)
    poses = sf.broadcast_poses([seq.T_C_R(i) for i in range(4)])
    for i in range(1, 4):
        sf.update(dev[i], poses[i])
    d, c = sf.gather_state()
    d = d.cpu().numpy()
    sf.close()
    d_ref, c_ref = np.full((h, w), 3.0), np.full((h, w), 3.0)
    for i in range(1, 4):
        T = seq.T_C_R(i)
        oracle.update(seq.params, frames[0], frames[i], T.q, T.t, d_ref, c_ref, rows=(20, 460), row_stride=8)
    assert depth_agreement(seq.params, d, d_ref, range(20, 460, 8), rtol=1e-6) > 0.9999


def test_deferred_fusion_equals_immediate_fusion(DF, seq640):
    """The fusion of update k normally runs inside update k+1's first kernel (advance_kernel, straight from registers).
    Flushing after every update (setup_kernel + fuse_kernel from the maps) and the debug-plane mode must give the same
    bits, the same counters, and match the oracle."""
    seq, frames = seq640
    p = seq.params
    h, w = seq.shape
    n = seq.n_frames - 1  # 5 updates
    res = []
    for mode in ("deferred", "flushed", "flags"):
        f = DF(p)
        f.set_reference(frames[0])
        f.fill_state(3.0, 3.0)
        f.enable_flags(mode == "flags")
        for i in range(1, n + 1):
            f.update(frames[i], seq.T_C_R(i))
            if mode == "flushed":
                f.flush()
        d, c = f.download_state()
        res.append((d, c, f.counters()))
        f.close()
    for d, c, cnt in res[1:]:
        assert np.array_equal(d, res[0][0], equal_nan=True) and np.array_equal(c, res[0][1], equal_nan=True)
        assert cnt == res[0][2]
    d_ref, c_ref = np.full((h, w), 3.0), np.full((h, w), 3.0)
    for i in range(1, n + 1):
        T = seq.T_C_R(i)
        oracle.update(p, frames[0], frames[i], T.q, T.t, d_ref, c_ref, rows=(20, 460), row_stride=16)
    assert depth_agreement(p, res[0][0], d_ref, range(20, 460, 16), rtol=1e-6) > 0.9999


def test_accessors_see_the_deferred_fusion(DF, seq640):
    """Every accessor runs a pending fusion first: counters, masks and a state re-load right after an update."""
    seq, frames = seq640
    p = seq.params
    h, w = seq.shape
    f = DF(p)
    f.set_reference(frames[0])
    f.fill_state(3.0, 3.0)
    for i in range(1, 4):
        f.update(frames[i], seq.T_C_R(i))
    cnt = f.counters()                      # no explicit flush before
    assert cnt["frames"] == 3 and cnt["accepted"] > 0 and cnt["active"] >= cnt["accepted"]
    f.update(frames[4], seq.T_C_R(4))
    mask = f.variance_mask(1.0)             # reads cov2: must include update 4
    d, c = f.download_state()
    assert np.array_equal(mask[20:-20, 20:-20] == 255, ~(c[20:-20, 20:-20] > 1.0))
    # replacing the state drops back to a full set-up from the maps: same result as a fresh filter
    f.update(frames[5], seq.T_C_R(5))
    f.upload_state(d, c)
    f.update(frames[5], seq.T_C_R(5))
    d1, c1 = f.download_state()
    f.close()
    g = DF(p)
    g.set_reference(frames[0])
    g.upload_state(d, c)
    g.update(frames[5], seq.T_C_R(5))
    d2, c2 = g.download_state()
    g.close()
    assert np.array_equal(d1, d2, equal_nan=True) and np.array_equal(c1, c2, equal_nan=True)


# ---- "next" rows against fixtures generated from the compiled reference ------------------------------
def test_next_rows_against_reference_fixtures(DF, seq640):
    """evaludateDepth ref:569-590, getMaskFromVariance ref:199-204 and getPointCloudFromImageAndDistance
    (utils/pointcloud/pointcloud_from_image_depth.h:42-89) on the device vs tests/golden/remode640_ref_next.npz, which
    tests/golden/make_golden.py produced by calling the reference's own functions; also through a block-cyclic split."""
    from parity import synthetic_state
    nxt = np.load(G / "remode640_ref_next.npz")
    seq, frames = seq640
    p = seq.params
    h, w = seq.shape
    depth, cov2, truth = synthetic_state(h, w)
    thr = float(nxt["max_variance"])
    sha = lambda a: hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()
    f = DF(p)
    f.upload_state(depth, cov2)
    f.set_truth(truth)
    s, n = f.evaluate_depth(thr)
    mask = f.variance_mask(thr)
    color = np.ascontiguousarray(np.stack([frames[0], 255 - frames[0], frames[0] // 2], axis=-1))
    xyz, rgb = f.point_cloud(color, thr)
    f.close()
    assert np.isclose((s / n) ** 0.5, float(nxt["rms"]), rtol=1e-10)
    assert sha(mask) == str(nxt["mask_sha"])
    assert len(xyz) == int(nxt["cloud_n"])
    assert sha(rgb) == str(nxt["cloud_rgb_sha"])
    assert np.array_equal(xyz[:64], nxt["cloud_xyz_head"]) and np.array_equal(xyz[::997], nxt["cloud_xyz_stride"])
    assert sha(xyz) == str(nxt["cloud_xyz_sha"])
    # block-cyclic contexts: each returns the points of its own rows, in scan order
    b = p.border
    valid = (depth != 0) & (mask == 255)
    per_row = np.zeros(h + 1, np.int64)
    per_row[b + 1:h - b + 1] = valid[b:h - b, b:w - b].sum(axis=1)
    off = np.cumsum(per_row)
    total = 0
    for part in range(2):
        g = DF(p, cyclic=(8, 2, part))
        g.upload_state(depth, cov2)
        x2, r2 = g.point_cloud(color, thr)
        rows = g.owned_rows()
        g.close()
        idx = np.concatenate([np.arange(off[y], off[y + 1]) for y in rows])
        assert np.array_equal(x2, xyz[idx]) and np.array_equal(r2, rgb[idx])
        total += len(x2)
    assert total == len(xyz)


def test_inverse_depth_variant_against_reference_fixture(DF):
    """USE_INVERSE_DEPTH_FOR_FILTERING 1 (ref:63): 4 updates from (3.0, 0.5) vs the variant translation unit's maps."""
    nxt = np.load(G / "remode640_ref_next.npz")
    seq = make_sequence("remode_640x480", n_frames=6, inverse_depth=True)
    p = seq.params
    frames = [seq.render_host(i) for i in range(5)]
    f = DF(p)
    f.set_reference(frames[0])
    f.fill_state(3.0, 0.5)
    for i in range(1, 5):
        f.update(frames[i], seq.T_C_R(i))
    d, c = f.download_state()
    f.close()
    st = int(nxt["inv_row_step"])
    rows = range(0, 480, st)
    d_ref, c_ref = _expand(nxt["inv_depth_rows"], st, d.shape), _expand(nxt["inv_cov2_rows"], st, c.shape)
    assert depth_agreement(p, d, d_ref, rows) >= MIN_DEPTH_AGREE
    assert depth_agreement(p, d, d_ref, rows, rtol=1e-6) > 0.999
    assert class_mismatch(p, c, c_ref, rows) <= MAX_DECISION_MISMATCH


def test_strict_dropin_caches_are_transparent(DF, seq640):
    """dmf_update_strict skips re-uploading an unchanged reference image / unchanged maps.  Neither cache may be
    observable: the caller edits the maps between calls, swaps the reference image for another one and back, and passes
    copies at new addresses — every call must equal the oracle's update() on the same arguments."""
    from slamplay_b200.depth_filter import release_strict_contexts, update
    seq, frames = seq640
    p = seq.params
    h, w = seq.shape
    depth, cov2 = np.full((h, w), 3.0), np.full((h, w), 3.0)
    d_ref, c_ref = depth.copy(), cov2.copy()
    plan = [(0, 1, None), (0, 2, None), (0, 3, "edit"), (1, 4, None), (0, 5, "copy"), (0, 2, "edit")]
    for ref_i, cur_i, action in plan:
        if action == "edit":      # the caller overwrites part of the state between two calls
            for m in (depth, d_ref):
                m[100:140, 200:300] = 2.5
            for m in (cov2, c_ref):
                m[100:140, 200:300] = 1.0
        if action == "copy":      # same contents at new addresses
            depth, cov2 = depth.copy(), cov2.copy()
        ref_img = frames[ref_i].copy() if action == "copy" else frames[ref_i]
        T = seq.T_C_R(cur_i)
        update(ref_img, frames[cur_i], T, depth, cov2)
        oracle.update(p, frames[ref_i], frames[cur_i], T.q, T.t, d_ref, c_ref)
        assert depth_agreement(p, depth, d_ref, rtol=1e-6) > 0.9999, (ref_i, cur_i, action)
        assert class_mismatch(p, cov2, c_ref) <= MAX_DECISION_MISMATCH
    release_strict_contexts()


def test_grouped_divisions_are_bit_identical_to_ieee_division():
    """dmf_geometry.h computes groups of quotients over one denominator with a shared reciprocal; on the device every
    quotient must carry the bits of __ddiv_rn (the host build of the same header, which the CPU suite pins to the
    oracle, uses plain IEEE division)."""
    from slamplay_b200 import _lib
    bad = C.c_uint64(123)
    assert _lib.load_dmf().dmf_selftest_division(0, 200_000_000, 20261017, C.byref(bad)) == 0
    assert bad.value == 0


def test_moment_table_kernels_agree_bulk_copy_and_legacy(DF, seq640):
    """moments_bulk_kernel (tiles staged in shared memory by cp.async.bulk + mbarrier) is used when the frame rows are
    16-byte aligned, moments_kernel otherwise: a device frame with a 4-byte aligned pitch must give the same bits."""
    import os

    import torch
    seq, frames = seq640
    p = seq.params
    h, w = seq.shape
    res = []
    os.environ["DMF_MOMENTS"] = "bulk"  # small frames default to the per-column kernel; read at context creation
    for pitch in (w, w + 4):  # 640: 16-byte aligned rows (bulk copies); 644: only 4-byte aligned (per-column kernel)
        dev = torch.zeros((3, h, pitch), dtype=torch.uint8, device="cuda")
        for i in range(3):
            dev[i, :, :w] = torch.from_numpy(frames[i + 1]).cuda()
        torch.cuda.synchronize()
        f = DF(p)
        f.set_reference(frames[0])
        f.fill_state(3.0, 3.0)
        for i in range(3):
            f.update_device(dev[i].data_ptr(), pitch, seq.T_C_R(i + 1))
        res.append(f.download_state() + (f.counters(),))
        f.close()
    os.environ.pop("DMF_MOMENTS", None)
    assert np.array_equal(res[0][0], res[1][0], equal_nan=True) and np.array_equal(res[0][1], res[1][1], equal_nan=True)
    assert res[0][2] == res[1][2]
