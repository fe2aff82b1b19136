"""Synthetic sequences shaped like BASELINE.json's configs (SURVEY.md §8d "Synthetic inputs").

A sequence = scene + intrinsics + a list of camera poses T_WC (frame 0 is the reference frame,
like the REMODE list read at dense_mapping/test_monocular_mapping.cpp:322-337) and is rendered
either on the CPU (libdmf_synth_cpu.so, for the CPU test-suite and golden fixtures) or on the
GPU (libdmf_synth.so, for the 1080p / 4K benchmark sequences).  Both renderers are bit-identical.
"""
from __future__ import annotations

import ctypes as C
import math
from dataclasses import dataclass, field
from typing import List, Optional, Tuple

import numpy as np

from . import _lib
from ._lib import DmfParams, SynthCamera, SynthScene
from .se3 import SE3, relative_pose

WORKLOADS = {
    # name: (width, height, frames incl. reference, kind)  — BASELINE.json configs
    "remode_640x480": (640, 480, 200, "lateral"),
    "kitti_1241x376": (1241, 376, 200, "forward"),
    "hd_1920x1080": (1920, 1080, 300, "lateral"),
    "uhd_3840x2160": (3840, 2160, 500, "lateral"),
}


def _rotvec_quat(rx: float, ry: float, rz: float) -> Tuple[float, float, float, float]:
    a = math.sqrt(rx * rx + ry * ry + rz * rz)
    if a < 1e-300:
        return (0.0, 0.0, 0.0, 1.0)
    s = math.sin(0.5 * a) / a
    return (rx * s, ry * s, rz * s, math.cos(0.5 * a))


@dataclass
class Sequence:
    name: str
    params: DmfParams
    scene: SynthScene
    poses_T_WC: List[SE3]
    kind: str = "lateral"
    _cam_cache: dict = field(default_factory=dict, repr=False)

    @property
    def n_frames(self) -> int:
        return len(self.poses_T_WC)

    @property
    def shape(self) -> Tuple[int, int]:
        return (self.params.height, self.params.width)

    def camera(self, i: int) -> SynthCamera:
        p = self.params
        T = self.poses_T_WC[i]
        cam = SynthCamera()
        cam.width, cam.height = p.width, p.height
        cam.fx, cam.fy, cam.cx, cam.cy = p.fx, p.fy, p.cx, p.cy
        cam.q = (C.c_double * 4)(*T.q)
        cam.t = (C.c_double * 3)(*T.t)
        return cam

    def T_C_R(self, i: int) -> SE3:
        """Pose handed to update() for frame i: T_WC(i)^-1 * T_WC(0)  (ref:289-290)."""
        return relative_pose(self.poses_T_WC[0], self.poses_T_WC[i])

    # -- CPU rendering -----------------------------------------------------------------
    def render_host(self, i: int, with_distance: bool = False):
        lib = _lib.load_synth_cpu()
        h, w = self.shape
        img = np.zeros((h, w), np.uint8)
        dist = np.zeros((h, w), np.float64) if with_distance else None
        cam = self.camera(i)
        rc = lib.dmf_synth_render_host(C.byref(self.scene), C.byref(cam), img.ctypes.data, img.strides[0],
                                       dist.ctypes.data if dist is not None else None,
                                       dist.strides[0] if dist is not None else 0)
        if rc != 0:
            raise RuntimeError("dmf_synth_render_host failed")
        return (img, dist) if with_distance else img

    # -- GPU rendering (into caller-provided device memory) ------------------------------
    def render_device(self, i: int, img_ptr: int, pitch: int, dist_ptr: int = 0, dist_pitch: int = 0,
                      stream: int = 0) -> None:
        lib = _lib.load_synth_cuda()
        cam = self.camera(i)
        rc = lib.dmf_synth_render_device(C.byref(self.scene), C.byref(cam), C.c_void_p(img_ptr), pitch,
                                         C.c_void_p(dist_ptr) if dist_ptr else None, dist_pitch,
                                         C.c_void_p(stream) if stream else None)
        if rc != 0:
            raise RuntimeError(f"dmf_synth_render_device failed ({rc})")


def make_params(width: int, height: int, kind: str = "lateral", inverse_depth: bool = False) -> DmfParams:
    """Reference constants (ref:72-89) with the intrinsics of the named shape.  Pure Python so the
    CPU test-suite does not need the CUDA library; identical to dmf_default_params for 'lateral'."""
    p = DmfParams()
    p.width, p.height, p.border, p.ncc_half = width, height, 20, 3
    f32 = lambda v: float(np.float32(v))
    if kind == "forward":  # KITTI-shaped intrinsics
        p.fx, p.fy, p.cx, p.cy = 718.856, 718.856, 607.19 * width / 1241.0, 185.22 * height / 376.0
    elif width == 640 and height == 480:
        p.fx, p.fy, p.cx, p.cy = f32(481.2), f32(-480.0), f32(319.5), f32(239.5)
    else:
        s = width / 640.0
        p.fx, p.fy, p.cx, p.cy = f32(481.2) * s, -480.0 * s, 0.5 * (width - 1), 0.5 * (height - 1)
    p.step, p.max_half_len, p.min_depth, p.n_sigma = 0.7, 100.0, 0.1, 3.0
    p.ncc_thresh = f32(0.85)
    if inverse_depth:
        p.min_cov, p.max_cov = 0.0001, 1.0
    else:
        good_error = 0.01
        p.min_cov, p.max_cov = good_error * good_error, 10.0
    p.inverse_depth = int(inverse_depth)
    return p


def make_sequence(name: str = "remode_640x480", n_frames: Optional[int] = None, seed: int = 0,
                  width: Optional[int] = None, height: Optional[int] = None, kind: Optional[str] = None,
                  inverse_depth: bool = False) -> Sequence:
    """Build one of the BASELINE.json-shaped sequences (or a custom size with width/height/kind)."""
    if name in WORKLOADS:
        w, h, n, k = WORKLOADS[name]
    else:
        if width is None or height is None:
            raise ValueError(f"unknown workload {name!r}; give width/height for a custom one")
        w, h, n, k = width, height, n_frames or 50, kind or "lateral"
    if width is not None:
        w = width
    if height is not None:
        h = height
    if kind is not None:
        k = kind
    if n_frames is not None:
        n = n_frames
    p = make_params(w, h, k, inverse_depth)

    scene = SynthScene()
    scene.seed = 0x5EED0000 + seed
    scene.ray_iters = 12
    scene.supersample = 2
    scene.tex_octaves = 5
    poses: List[SE3] = []
    if k == "lateral":
        # REMODE-shaped: camera ~2 m from the surface, sweeping sideways with a slow wobble.
        scene.plane_z, scene.relief_amp, scene.relief_period = 2.0, 0.15, 1.0
        gsd = scene.plane_z / abs(p.fx)             # metres per pixel on the surface
        scene.tex_base = 1.5 * gsd
        step = 0.004 * 640.0 / w                    # keeps the per-frame disparity (in px) resolution independent
        for i in range(n):
            q = _rotvec_quat(0.02 * math.sin(0.05 * i), 0.02 * math.sin(0.04 * i + 1.0), 0.01 * math.sin(0.03 * i))
            t = (step * i, 0.25 * step * i * math.sin(0.11 * i), 0.05 * math.sin(0.03 * i))
            poses.append(SE3.from_quat_trans(q[0], q[1], q[2], q[3], *t))
    elif k == "forward":
        # KITTI-shaped: motion along the optical axis towards a relief wall ~5 m ahead.
        scene.plane_z, scene.relief_amp, scene.relief_period = 5.0, 0.5, 3.0
        gsd = scene.plane_z / abs(p.fx)
        scene.tex_base = 1.5 * gsd
        for i in range(n):
            q = _rotvec_quat(0.004 * math.sin(0.05 * i), 0.004 * math.sin(0.04 * i + 1.0), 0.002 * math.sin(0.03 * i))
            t = (0.002 * i, 0.0005 * i, 0.010 * i)
            poses.append(SE3.from_quat_trans(q[0], q[1], q[2], q[3], *t))
    else:
        raise ValueError(f"unknown sequence kind {k!r}")
    return Sequence(name=name, params=p, scene=scene, poses_T_WC=poses, kind=k)
