"""Generates tests/golden/*.npz from the COMPILED REFERENCE translation unit (oracle/_ref/libdmf_ref.so,
built by oracle/Makefile from /root/reference/dense_mapping/test_monocular_mapping.cpp against the
stand-in third-party headers).  Run in the container that has /root/reference:

    python tests/golden/make_golden.py

The fixtures pin the oracle (tests/test_oracle_golden.py, CPU) and the CUDA path
(tests/test_gpu_parity.py::test_golden_sequence, GPU) to outputs of the reference's own code.
Inputs are regenerated from the deterministic synthetic renderer; their SHA-256 is stored so that a
drift of the generator is detected rather than silently changing the question.
"""
import ctypes as C
import hashlib
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
import oracle  # noqa: E402
from slamplay_b200.synth import make_sequence  # noqa: E402
from parity import synthetic_state  # noqa: E402

HERE = Path(__file__).resolve().parent
N_FRAMES = 6
ROW_STEP = 8  # store every 8th row of the final maps (plus SHA-256 of the full maps)


def sha(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def main():
    oracle.build(ref=True)
    L = oracle.ref_lib()
    assert L is not None, "oracle/_ref/libdmf_ref.so missing: needs /root/reference"
    seq = make_sequence("remode_640x480", n_frames=N_FRAMES)
    frames = [seq.render_host(i) for i in range(N_FRAMES)]
    h, w = seq.shape
    depth = np.full((h, w), 3.0)
    cov2 = np.full((h, w), 3.0)
    poses = []
    per_frame = []
    for i in range(1, N_FRAMES):
        T = seq.T_C_R(i)
        poses.append(list(T.q) + list(T.t))
        oracle.ref_update(frames[0], frames[i], T.q, T.t, depth, cov2)
        per_frame.append([sha(depth), sha(cov2)])
    np.savez_compressed(
        HERE / "remode640_ref_update.npz",
        n_frames=N_FRAMES, poses=np.array(poses), frame_sha=np.array([sha(f) for f in frames]),
        depth_rows=depth[::ROW_STEP].copy(), cov2_rows=cov2[::ROW_STEP].copy(), row_step=ROW_STEP,
        depth_sha=sha(depth), cov2_sha=sha(cov2), per_frame_sha=np.array(per_frame))

    # unit-level vectors from the reference's own NCC / epipolarSearch / updateDepthFilter
    rng = np.random.default_rng(1234)
    ref, cur = frames[0], frames[3]
    T = seq.T_C_R(3)
    q, t = (C.c_double * 4)(*T.q), (C.c_double * 3)(*T.t)
    n = 256
    rx = rng.integers(20, w - 20, n).astype(np.float64)
    ry = rng.integers(20, h - 20, n).astype(np.float64)
    cx = rng.uniform(20, w - 21, n)
    cy = rng.uniform(20, h - 20, n)
    ncc = np.array([L.ref_ncc(ref.ctypes.data, ref.strides[0], cur.ctypes.data, cur.strides[0], rx[i], ry[i], cx[i], cy[i])
                    for i in range(n)])
    mu = rng.uniform(1.0, 4.0, n)
    sigma = rng.uniform(0.02, 1.5, n)
    es = np.zeros((n, 5))
    for i in range(n):
        out = (C.c_double * 5)()
        L.ref_epipolar_search(ref.ctypes.data, ref.strides[0], cur.ctypes.data, cur.strides[0], q, t, rx[i], ry[i], mu[i], sigma[i], out)
        es[i] = list(out)
    fu = np.zeros((n, 2))
    dirs = rng.normal(size=(n, 2))
    dirs /= np.linalg.norm(dirs, axis=1, keepdims=True)
    dval = rng.uniform(1.0, 4.0, n)
    cval = rng.uniform(1e-3, 3.0, n)
    for i in range(n):
        out = (C.c_double * 2)()
        L.ref_update_depth_filter(q, t, int(rx[i]), int(ry[i]), cx[i], cy[i], dirs[i, 0], dirs[i, 1], dval[i], cval[i], out)
        fu[i] = list(out)
    np.savez_compressed(HERE / "remode640_ref_units.npz", frame_ref=0, frame_cur=3, pose=np.array(list(T.q) + list(T.t)),
                        rx=rx, ry=ry, cx=cx, cy=cy, ncc=ncc, mu=mu, sigma=sigma, search=es, dirs=dirs, dval=dval, cval=cval, fuse=fu)
    make_next_rows(seq, frames)
    print("wrote", [p.name for p in HERE.glob("*.npz")])


def make_next_rows(seq, frames):
    """Fixtures of the 'next' rows from the compiled reference: evaludateDepth ref:569-590,
    getMaskFromVariance ref:199-204, getPointCloudFromImageAndDistance (the reference's own header), the inverse-depth
    variant (ref:63, libdmf_ref_inv.so), readDatasetFiles ref:317-352 and the pose chain ref:289-290."""
    import tempfile

    from slamplay_b200.remode import POSE_FILE, write_dataset
    h, w = seq.shape
    depth, cov2, truth = synthetic_state(h, w)
    thr = 2e-4  # good_cov ref:89
    rms = oracle.ref_evaluate_depth(truth, depth, cov2, thr)
    mask = oracle.ref_variance_mask(cov2, thr)
    color = np.ascontiguousarray(np.stack([frames[0], 255 - frames[0], frames[0] // 2], axis=-1))
    xyz, rgb = oracle.ref_point_cloud(color, depth, mask)
    # inverse-depth variant: 4 updates from (3.0, 0.5)
    seqi = make_sequence("remode_640x480", n_frames=N_FRAMES, inverse_depth=True)
    di, ci = np.full((h, w), 3.0), np.full((h, w), 0.5)
    inv_sha = []
    for i in range(1, 5):
        T = seqi.T_C_R(i)
        oracle.ref_update(frames[0], frames[i], T.q, T.t, di, ci, inverse=True)
        inv_sha.append([sha(di), sha(ci)])
    # reader: a 3-frame REMODE-layout directory written by our writer, read by the reference's readDatasetFiles
    with tempfile.TemporaryDirectory() as tmp:
        _, gt = seq.render_host(0, with_distance=True)
        write_dataset(tmp, seq, frames[:3], gt)
        files, poses, ref_depth = oracle.ref_read_dataset(tmp)
        pose_txt = open(Path(tmp) / POSE_FILE).read()
    T_C_R = np.array([oracle.ref_compose_T_C_R(poses[0], poses[k]) for k in range(3)])
    np.savez_compressed(
        HERE / "remode640_ref_next.npz", max_variance=thr, rms=rms, mask_sha=sha(mask), mask_rows=mask[::32].copy(),
        cloud_n=len(xyz), cloud_xyz_sha=sha(xyz), cloud_rgb_sha=sha(rgb), cloud_xyz_head=xyz[:64].copy(), cloud_rgb_head=rgb[:64].copy(),
        cloud_xyz_stride=xyz[::997].copy(),
        inv_per_frame_sha=np.array(inv_sha), inv_depth_rows=di[::16].copy(), inv_cov2_rows=ci[::16].copy(), inv_row_step=16,
        reader_pose_txt=pose_txt, reader_poses=poses[:3].copy(), reader_n_entries=len(files), reader_T_C_R=T_C_R,
        reader_depth_sha=sha(ref_depth), state_sha=np.array([sha(depth), sha(cov2), sha(truth)]))


if __name__ == "__main__":
    main()
