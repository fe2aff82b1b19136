"""slamplay_b200 — B200-native dense monocular depth filter.

A from-scratch sm_100a implementation of ONE hot path of luigifreda/slamplay: the per-pixel
`update()` loop of dense_mapping/test_monocular_mapping.cpp (epipolar NCC search + depth-filter
fusion), behind the reference's own call surface.  See DESIGN.md.
"""
from .se3 import SE3, relative_pose  # noqa: F401

__all__ = ["SE3", "relative_pose", "DepthFilter", "update", "default_params", "DmfError"]


def __getattr__(name):  # lazy: importing the package must not require the CUDA library
    if name in ("DepthFilter", "update", "default_params", "DmfError", "release_strict_contexts"):
        from . import depth_filter
        return getattr(depth_filter, name)
    raise AttributeError(name)
