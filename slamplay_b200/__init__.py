"""slamplay_b200 — B200-native dense monocular depth filter.

A from-scratch sm_100a implementation of ONE hot path of luigifreda/slamplay: the per-pixel
`update()` loop of dense_mapping/test_monocular_mapping.cpp (epipolar NCC search + depth-filter
fusion), behind the reference's own call surface.  See DESIGN.md.
"""
import os as _os

# The frame ring (frame_ring.py) orders streams of different processes with stream memory operations: a stream that waits
# on a flag must not share a hardware queue with the stream that will set it.  Effective when set before CUDA initialises.
_os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

from .se3 import SE3, relative_pose  # noqa: E402,F401

__all__ = ["SE3", "relative_pose", "DepthFilter", "update", "default_params", "DmfError"]


def __getattr__(name):  # lazy: importing the package must not require the CUDA library
    if name in ("DepthFilter", "update", "default_params", "DmfError", "release_strict_contexts"):
        from . import depth_filter
        return getattr(depth_filter, name)
    raise AttributeError(name)
