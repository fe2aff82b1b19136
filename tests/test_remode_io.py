"""REMODE on-disk format (ref:317-352): writer/reader round trip and the driver's pose chain."""
import numpy as np

from slamplay_b200.remode import read_dataset, write_dataset
from slamplay_b200.se3 import relative_pose
from slamplay_b200.synth import make_sequence


def test_round_trip(tmp_path):
    import cv2
    seq = make_sequence("tiny", width=160, height=120, n_frames=3)
    frames = [seq.render_host(i) for i in range(3)]
    _, gt = seq.render_host(0, with_distance=True)
    write_dataset(str(tmp_path), seq, frames, gt)
    files, poses, depth = read_dataset(str(tmp_path), 160, 120)
    assert len(files) == 3 and files[1].endswith("images/scene_001.png")
    for a, b in zip(poses, seq.poses_T_WC):
        assert np.allclose(a.q, b.q, atol=1e-15) and a.t == b.t  # repr() round-trips doubles; q is re-normalised
    assert np.allclose(depth, gt, rtol=1e-15)                       # /100 after *100
    img = cv2.imread(files[2], cv2.IMREAD_GRAYSCALE)                # imread(..., 0) ref:287
    assert np.array_equal(img, frames[2])
    T = relative_pose(poses[0], poses[2])                           # ref:289-290
    assert np.allclose(T.t, seq.T_C_R(2).t, atol=1e-15)
