"""REMODE test-data reader — the step before the path.

Mirrors readDatasetFiles() of dense_mapping/test_monocular_mapping.cpp:317-352:

* `<dir>/first_200_frames_traj_over_table_input_sequence.txt`: one line per frame,
  `image tx ty tz qx qy qz qw` with the pose T_WC (ref:326); poses become
  SE3d(Quaterniond(qw,qx,qy,qz), Vector3d(tx,ty,tz)) (ref:333-335, the constructor normalises q);
  images live in `<dir>/images/<image>` (ref:332).
* `<dir>/depthmaps/scene_000.depth`: width*height numbers, row-major, in centimetres — the
  reference divides by 100 (ref:341-349).

`write_dataset` produces the same layout from a synthetic sequence, so the reader (and the whole driver
loop ref:264-305) can be exercised without the real download (scripts/download_dataset_remode_test_data.sh).
"""
from __future__ import annotations

import os
from pathlib import Path
from typing import List, Tuple

import numpy as np

from .se3 import SE3

POSE_FILE = "first_200_frames_traj_over_table_input_sequence.txt"
DEPTH_FILE = "depthmaps/scene_000.depth"


def read_poses(path: str) -> Tuple[List[str], List[SE3]]:
    """The pose list (ref:322-337): (image file paths, poses T_WC).  Complete entries only: the reference's
    `while (!fin.eof())` loop appends one bogus entry after a trailing newline, which its driver skips (ref:288)."""
    root = Path(path)
    files: List[str] = []
    poses: List[SE3] = []
    with open(root / POSE_FILE) as f:
        tokens = f.read().split()
    # the reference reads token-wise with operator>> until EOF (ref:325-337)
    for i in range(0, len(tokens) - 7, 8):
        name = tokens[i]
        tx, ty, tz, qx, qy, qz, qw = (float(v) for v in tokens[i + 1:i + 8])
        files.append(str(root) + "/images/" + name)  # path + "/images/" + image, ref:332
        poses.append(SE3.from_quat_trans(qx, qy, qz, qw, tx, ty, tz))
    return files, poses


def read_dataset(path: str, width: int = 640, height: int = 480) -> Tuple[List[str], List[SE3], np.ndarray]:
    """Returns (image file paths, poses T_WC, reference depth map in metres (H, W) float64)."""
    root = Path(path)
    files, poses = read_poses(path)
    depth = np.loadtxt(root / DEPTH_FILE, dtype=np.float64).reshape(-1)
    if depth.size != width * height:
        raise ValueError(f"{DEPTH_FILE}: expected {width * height} values, found {depth.size}")
    return files, poses, depth.reshape(height, width) / 100.0


def write_dataset(path: str, seq, frames, ref_distance: np.ndarray, ext: str = "png") -> None:
    """Writes a synthetic sequence in the REMODE layout (images through cv2: "png" like the real set, or "pgm" for readers
    without an image library, e.g. slamplay_b200/cpp/example_remode_dir.cpp)."""
    import cv2

    root = Path(path)
    os.makedirs(root / "images", exist_ok=True)
    os.makedirs(root / "depthmaps", exist_ok=True)
    with open(root / POSE_FILE, "w") as f:
        for i, (img, T) in enumerate(zip(frames, seq.poses_T_WC)):
            name = f"scene_{i:03d}.{ext}"
            cv2.imwrite(str(root / "images" / name), img)
            q, t = T.q, T.t
            f.write(f"{name} {t[0]!r} {t[1]!r} {t[2]!r} {q[0]!r} {q[1]!r} {q[2]!r} {q[3]!r}\n")
    np.savetxt(root / DEPTH_FILE, (ref_distance * 100.0).reshape(-1), fmt="%.17g")
