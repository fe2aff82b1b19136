"""Host-side mirror of the reference's call surface for the dense monocular depth filter.

Reference: luigifreda/slamplay dense_mapping/test_monocular_mapping.cpp

    void update(const Mat &ref, const Mat &curr, const SE3d &T_C_R, Mat &depth, Mat &depth_cov2);   # :107-112, :355

* `update(ref, curr, T_C_R, depth, depth_cov2)` below has the same name, argument order and
  in-place semantics (numpy arrays stand in for cv::Mat, `SE3` for Sophus::SE3d).  It is the
  STRICT drop-in: maps are uploaded, updated on the GPU and downloaded before it returns, as the
  reference's caller reads them after every call (:292-300).
* `DepthFilter` is the RESIDENT form north_star describes: depth / depth_cov2 stay in HBM for the
  whole sequence; only the u8 frame and the 7-double pose cross PCIe per update.

Everything computes through the C ABI of include/dmf.h (libdmf.so, sm_100a).  There is no
CPU fallback: without the CUDA library or a B200 these calls raise.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence, Tuple

import numpy as np

from . import _lib
from ._lib import DmfCounters, DmfParams
from .se3 import SE3


class DmfError(RuntimeError):
    pass


def default_params(width: int = 640, height: int = 480, inverse_depth: bool = False) -> DmfParams:
    """Reference constants (:72-89) for a width x height image; see dmf_default_params."""
    p = DmfParams()
    rc = _lib.load_dmf().dmf_default_params(C.byref(p), int(width), int(height), int(bool(inverse_depth)))
    if rc != 0:
        raise DmfError(_lib.load_dmf().dmf_last_error(None).decode())
    return p


def _pose_arrays(T_C_R) -> Tuple[C.Array, C.Array]:
    if isinstance(T_C_R, SE3):
        q, t = T_C_R.q, T_C_R.t
    else:
        q, t = T_C_R
    if len(q) != 4 or len(t) != 3:
        raise ValueError("T_C_R must be an SE3 or a (q_xyzw[4], t[3]) pair")
    return (C.c_double * 4)(*[float(v) for v in q]), (C.c_double * 3)(*[float(v) for v in t])


def _check_u8(name: str, img: np.ndarray, p: DmfParams) -> np.ndarray:
    if not isinstance(img, np.ndarray) or img.dtype != np.uint8 or img.ndim != 2:
        raise ValueError(f"{name} must be a 2-D uint8 array (CV_8UC1)")
    if img.shape != (p.height, p.width):
        raise ValueError(f"{name} has shape {img.shape}, expected {(p.height, p.width)}")  # MSG_ASSERT at :265
    if img.strides[1] != 1:
        raise ValueError(f"{name} must have contiguous rows")
    return img


def _check_f64(name: str, m: np.ndarray, p: DmfParams) -> np.ndarray:
    if not isinstance(m, np.ndarray) or m.dtype != np.float64 or m.ndim != 2:
        raise ValueError(f"{name} must be a 2-D float64 array (CV_64F)")
    if m.shape != (p.height, p.width):
        raise ValueError(f"{name} has shape {m.shape}, expected {(p.height, p.width)}")
    if m.strides[1] != 8:
        raise ValueError(f"{name} must have contiguous rows")
    return m


class DepthFilter:
    """Resident depth-filter context (one per GPU / row band)."""

    def __init__(self, params: Optional[DmfParams] = None, *, width: int = 640, height: int = 480,
                 device: int = 0, rows: Optional[Tuple[int, int]] = None, inverse_depth: bool = False,
                 cyclic: Optional[Tuple[int, int, int]] = None):
        """rows=(r0, r1): contiguous band; cyclic=(block_rows, n_parts, part): block-cyclic row ownership."""
        self._lib = _lib.load_dmf()
        self.params = params.copy() if params is not None else default_params(width, height, inverse_depth)
        r0, r1 = rows if rows is not None else (0, self.params.height)
        self._ctx = C.c_void_p()
        if cyclic is not None:
            rc = self._lib.dmf_create_cyclic(C.byref(self.params), int(device), int(cyclic[0]), int(cyclic[1]), int(cyclic[2]),
                                             C.byref(self._ctx))
        else:
            rc = self._lib.dmf_create(C.byref(self.params), int(device), int(r0), int(r1), C.byref(self._ctx))
        if rc != 0:
            msg = self._lib.dmf_last_error(None).decode()
            self._ctx = C.c_void_p()
            raise DmfError(f"dmf_create failed ({rc}): {msg}")
        self.device = int(device)
        self.rows = (int(r0), int(r1))
        self._keep = []  # pinned/host frames that must outlive async copies

    # -- lifecycle ---------------------------------------------------------------------
    def close(self) -> None:
        if getattr(self, "_ctx", None) is not None and self._ctx.value:
            self._lib.dmf_destroy(self._ctx)
            self._ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def _ck(self, rc: int, what: str) -> None:
        if rc != 0:
            raise DmfError(f"{what} failed ({rc}): {self._lib.dmf_last_error(self._ctx).decode()}")

    # -- inputs ------------------------------------------------------------------------
    def set_reference(self, ref: np.ndarray) -> None:
        ref = _check_u8("ref", ref, self.params)
        self._ck(self._lib.dmf_set_reference(self._ctx, ref.ctypes.data, ref.strides[0]), "dmf_set_reference")

    def set_reference_device(self, ptr: int, step: int) -> None:
        self._ck(self._lib.dmf_set_reference_device(self._ctx, C.c_void_p(ptr), step), "dmf_set_reference_device")

    def fill_state(self, init_depth: float = 3.0, init_cov2: float = 3.0) -> None:
        """Mat(h, w, CV_64F, init) of :277-278 (init_depth = init_cov2 = 3.0 at :270,274)."""
        if not init_cov2 < self.params.max_cov:
            raise ValueError("Please increase max_cov above the init cov")  # MSG_ASSERT at :276
        self._ck(self._lib.dmf_fill_state(self._ctx, float(init_depth), float(init_cov2)), "dmf_fill_state")

    def upload_state(self, depth: np.ndarray, depth_cov2: np.ndarray) -> None:
        d = _check_f64("depth", depth, self.params)
        c = _check_f64("depth_cov2", depth_cov2, self.params)
        self._ck(self._lib.dmf_upload_state(self._ctx, d.ctypes.data, d.strides[0], c.ctypes.data, c.strides[0]),
                 "dmf_upload_state")

    def download_state(self, depth: Optional[np.ndarray] = None, depth_cov2: Optional[np.ndarray] = None):
        p = self.params
        if depth is None:
            depth = np.zeros((p.height, p.width), np.float64)
        if depth_cov2 is None:
            depth_cov2 = np.zeros((p.height, p.width), np.float64)
        d = _check_f64("depth", depth, p)
        c = _check_f64("depth_cov2", depth_cov2, p)
        self._ck(self._lib.dmf_download_state(self._ctx, d.ctypes.data, d.strides[0], c.ctypes.data, c.strides[0]),
                 "dmf_download_state")
        return depth, depth_cov2

    # -- the hot path --------------------------------------------------------------------
    def update(self, curr: np.ndarray, T_C_R) -> None:
        """One reference update() (:355-393) against `curr`; asynchronous."""
        curr = _check_u8("curr", curr, self.params)
        q, t = _pose_arrays(T_C_R)
        self._ck(self._lib.dmf_update(self._ctx, curr.ctypes.data, curr.strides[0], q, t), "dmf_update")

    def update_ptr(self, host_ptr: int, step: int, T_C_R) -> None:
        q, t = _pose_arrays(T_C_R)
        self._ck(self._lib.dmf_update(self._ctx, C.c_void_p(host_ptr), step, q, t), "dmf_update")

    def update_device(self, dev_ptr: int, step: int, T_C_R, wait_stream: Optional[int] = None) -> None:
        q, t = _pose_arrays(T_C_R)
        self._ck(self._lib.dmf_update_device(self._ctx, C.c_void_p(dev_ptr), step, q, t,
                                             C.c_void_p(wait_stream) if wait_stream else None), "dmf_update_device")

    def update_strict(self, ref: np.ndarray, curr: np.ndarray, T_C_R, depth: np.ndarray, depth_cov2: np.ndarray) -> None:
        """The reference's update(ref, curr, T_C_R, depth, depth_cov2) in one call (dmf_update_strict): the maps are valid
        in host memory on return; the reference image and unchanged maps are not uploaded again."""
        p = self.params
        ref, curr = _check_u8("ref", ref, p), _check_u8("curr", curr, p)
        d, c = _check_f64("depth", depth, p), _check_f64("depth_cov2", depth_cov2, p)
        q, t = _pose_arrays(T_C_R)
        self._ck(self._lib.dmf_update_strict(self._ctx, ref.ctypes.data, ref.strides[0], curr.ctypes.data, curr.strides[0], q, t,
                                             d.ctypes.data, d.strides[0], c.ctypes.data, c.strides[0]), "dmf_update_strict")

    def update_ring(self, ring, T_C_R) -> None:
        """update() against the next frame of a FrameRing (dmf_update_ring): the frame is pulled into this context's
        buffer by a copy engine, ordered against the producer by stream memory operations; asynchronous."""
        q, t = _pose_arrays(T_C_R)
        self._ck(self._lib.dmf_update_ring(self._ctx, ring._ptr, q, t), "dmf_update_ring")

    def flush(self) -> None:
        """Enqueue the deferred fusion of the last update (asynchronous); every accessor does this implicitly."""
        self._ck(self._lib.dmf_flush(self._ctx), "dmf_flush")

    def sync(self) -> None:
        self._ck(self._lib.dmf_sync(self._ctx), "dmf_sync")

    # -- introspection -------------------------------------------------------------------
    def counters(self, reset: bool = False) -> dict:
        out = DmfCounters()
        self._ck(self._lib.dmf_read_counters(self._ctx, C.byref(out), int(reset)), "dmf_read_counters")
        return out.as_dict()

    def set_timing(self, on: bool = True) -> None:
        self._ck(self._lib.dmf_set_timing(self._ctx, int(on)), "dmf_set_timing")

    def timing(self, reset: bool = False) -> dict:
        """Accumulated per-kernel milliseconds and the frames they cover: setup_ms = advance_kernel (fusion of the previous
        update + set-up) or setup_kernel, moments_ms, ncc_ms, fuse_ms = stand-alone fusion behind the update (debug planes)."""
        ms = (C.c_double * 4)()
        n = C.c_uint64()
        self._ck(self._lib.dmf_get_timing(self._ctx, ms, C.byref(n), int(reset)), "dmf_get_timing")
        return {"setup_ms": ms[0], "moments_ms": ms[1], "ncc_ms": ms[2], "fuse_ms": ms[3], "frames": int(n.value)}  # launch order

    def enable_flags(self, on: bool = True) -> None:
        self._ck(self._lib.dmf_enable_flags(self._ctx, int(on)), "dmf_enable_flags")

    def flags(self) -> np.ndarray:
        p = self.params
        f = np.zeros((p.height, p.width), np.uint8)
        self._ck(self._lib.dmf_download_flags(self._ctx, f.ctypes.data, f.strides[0]), "dmf_download_flags")
        return f

    def debug(self):
        """(best NCC float32 HxW, trip count HxW, winning iteration HxW [-1: none]) of the last update."""
        p = self.params
        ncc = np.zeros((p.height, p.width), np.float32)
        raw = np.zeros((p.height, p.width), np.int32)
        self._ck(self._lib.dmf_download_debug(self._ctx, ncc.ctypes.data, raw.ctypes.data), "dmf_download_debug")
        best = raw & 0xFFFF
        return ncc, raw >> 16, np.where(best == 0xFFFF, -1, best)

    def band(self) -> Tuple[int, int]:
        a, b = C.c_int(), C.c_int()
        self._ck(self._lib.dmf_get_band(self._ctx, C.byref(a), C.byref(b)), "dmf_get_band")
        return a.value, b.value

    def owned_rows(self) -> np.ndarray:
        """Image rows this context updates, in local order."""
        n = C.c_int()
        self._ck(self._lib.dmf_get_rows(self._ctx, None, 0, C.byref(n)), "dmf_get_rows")
        out = (C.c_int * max(n.value, 1))()
        self._ck(self._lib.dmf_get_rows(self._ctx, out, n.value, C.byref(n)), "dmf_get_rows")
        return np.array(out[: n.value], dtype=np.int64)

    def device_state(self) -> Tuple[int, int, int]:
        d, c, pitch = C.c_void_p(), C.c_void_p(), C.c_size_t()
        self._ck(self._lib.dmf_device_state(self._ctx, C.byref(d), C.byref(c), C.byref(pitch)), "dmf_device_state")
        return d.value, c.value, pitch.value

    def stream(self) -> int:
        s = C.c_void_p()
        self._ck(self._lib.dmf_stream(self._ctx, C.byref(s)), "dmf_stream")
        return s.value or 0

    # -- "next" rows -------------------------------------------------------
    def set_truth(self, truth: np.ndarray) -> None:
        t = _check_f64("truth", truth, self.params)
        self._ck(self._lib.dmf_set_truth(self._ctx, t.ctypes.data, t.strides[0]), "dmf_set_truth")

    def evaluate_depth(self, max_variance: Optional[float] = None) -> Tuple[float, int]:
        """evaludateDepth (:569-590): returns (sum of squared errors, count); RMS = sqrt(s/n)."""
        if max_variance is None:
            max_variance = 2.0 * self.params.min_cov  # good_cov :89
        s, n = C.c_double(), C.c_uint64()
        self._ck(self._lib.dmf_evaluate_depth(self._ctx, float(max_variance), C.byref(s), C.byref(n)), "dmf_evaluate_depth")
        return s.value, n.value

    def variance_mask(self, max_variance: Optional[float] = None) -> np.ndarray:
        """getMaskFromVariance (:199-204) for the band rows."""
        if max_variance is None:
            max_variance = 2.0 * self.params.min_cov
        p = self.params
        m = np.zeros((p.height, p.width), np.uint8)
        self._ck(self._lib.dmf_variance_mask(self._ctx, float(max_variance), m.ctypes.data, m.strides[0]), "dmf_variance_mask")
        return m


    def point_cloud(self, color: np.ndarray, max_variance: Optional[float] = None):
        """getPointCloudFromImageAndDistance (utils/pointcloud/pointcloud_from_image_depth.h:42-89) as called at
        ref:296-300.  color: (H, W, 3) BGR or (H, W) gray uint8.  Returns (xyz float32 (N,3), rgb uint8 (N,3))."""
        if max_variance is None:
            max_variance = 2.0 * self.params.min_cov
        p = self.params
        if color.dtype != np.uint8 or color.shape[:2] != (p.height, p.width):
            raise ValueError("color must be a uint8 image of the filter's size")
        channels = 1 if color.ndim == 2 else color.shape[2]
        color = np.ascontiguousarray(color)
        cap = (p.height - 2 * p.border) * (p.width - 2 * p.border)
        xyz = np.zeros((cap, 3), np.float32)
        rgb = np.zeros((cap, 3), np.uint8)
        n = C.c_uint64()
        self._ck(self._lib.dmf_point_cloud(self._ctx, color.ctypes.data, color.strides[0], channels, float(max_variance),
                                           xyz.ctypes.data, rgb.ctypes.data, cap, C.byref(n)), "dmf_point_cloud")
        return xyz[: n.value], rgb[: n.value]


# ---------------------------------------------------------------------------------------------
# Strict drop-in: the reference's free function.
_strict_ctx: dict = {}


def update(ref: np.ndarray, curr: np.ndarray, T_C_R, depth: np.ndarray, depth_cov2: np.ndarray,
           params: Optional[DmfParams] = None, device: int = 0) -> None:
    """void update(const Mat &ref, const Mat &curr, const SE3d &T_C_R, Mat &depth, Mat &depth_cov2)

    Same contract as dense_mapping/test_monocular_mapping.cpp:355-393: `depth` and `depth_cov2`
    (float64, H x W) are updated in place and are valid on return.  `params` defaults to the
    reference constants for the image size.  A context per (size, params, device) is cached.
    """
    if params is None:
        if not isinstance(ref, np.ndarray) or ref.ndim != 2:
            raise ValueError("ref must be a 2-D uint8 array (CV_8UC1)")
        params = default_params(ref.shape[1], ref.shape[0])
    key = (bytes(params), int(device))
    f = _strict_ctx.get(key)
    if f is None:
        f = DepthFilter(params, device=device)
        _strict_ctx[key] = f
    f.update_strict(ref, curr, T_C_R, depth, depth_cov2)


def release_strict_contexts() -> None:
    for f in _strict_ctx.values():
        f.close()
    _strict_ctx.clear()
