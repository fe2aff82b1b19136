"""Python handle on the CPU oracle (TEST INFRASTRUCTURE ONLY).

May be imported only by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs.  Nothing under slamplay_b200/ imports this package.

`liboracle.so`       — the from-scratch FP64 restatement (oracle/dense_mono_oracle.cpp).
`_ref/libdmf_ref.so` — the UNMODIFIED reference translation unit compiled against stand-in
                       third-party headers (oracle/ref_shim/); 640x480 only; built only where
                       /root/reference exists (this container), shipped prebuilt to the GPU box.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path
from typing import Optional

import numpy as np

HERE = Path(__file__).resolve().parent
REF_ROOT = Path(os.environ.get("DMF_REF_ROOT", "/root/reference"))

_vp = C.c_void_p
_P = C.POINTER


class Params(C.Structure):  # struct dmf_params, include/dmf.h
    _fields_ = [
        ("width", C.c_int32), ("height", C.c_int32), ("border", C.c_int32), ("ncc_half", C.c_int32),
        ("fx", C.c_double), ("fy", C.c_double), ("cx", C.c_double), ("cy", C.c_double),
        ("step", C.c_double), ("max_half_len", C.c_double), ("min_depth", C.c_double), ("n_sigma", C.c_double),
        ("ncc_thresh", C.c_double), ("min_cov", C.c_double), ("max_cov", C.c_double),
        ("inverse_depth", C.c_int32), ("reserved", C.c_int32),
    ]


class Counters(C.Structure):
    _fields_ = [("frames", C.c_uint64), ("interior", C.c_uint64), ("active", C.c_uint64),
                ("ncc_evals", C.c_uint64), ("accepted", C.c_uint64)]

    def as_dict(self):
        return {k: int(getattr(self, k)) for k, _ in self._fields_}


def build(ref: bool = True, quiet: bool = True) -> None:
    """make liboracle.so and, when the reference tree is present, _ref/libdmf_ref.so."""
    out = subprocess.DEVNULL if quiet else None
    subprocess.run(["make", "-C", str(HERE), "all"], check=True, stdout=out)
    if ref and (REF_ROOT / "dense_mapping" / "test_monocular_mapping.cpp").exists():
        subprocess.run(["make", "-C", str(HERE), "ref", f"REF_ROOT={REF_ROOT}"], check=True, stdout=out)


_libs: dict = {}


def lib() -> C.CDLL:
    if "o" not in _libs:
        path = HERE / "liboracle.so"
        if not path.exists():
            build(ref=False)
        L = C.CDLL(str(path))
        L.dmo_default_params.argtypes = [_P(Params), C.c_int, C.c_int, C.c_int]
        L.dmo_update.argtypes = [_P(Params), _vp, C.c_size_t, _vp, C.c_size_t, _P(C.c_double), _P(C.c_double),
                                 _vp, C.c_size_t, _vp, C.c_size_t, C.c_int, C.c_int, C.c_int, C.c_int,
                                 _P(Counters), _vp, C.c_size_t, _vp, _vp]
        L.dmo_update_ex.argtypes = [_P(Params), _vp, C.c_size_t, _vp, C.c_size_t, _P(C.c_double), _P(C.c_double),
                                    _vp, C.c_size_t, _vp, C.c_size_t, C.c_int, C.c_int, C.c_int,
                                    _P(Counters), _vp, C.c_size_t, _vp, _vp]
        L.dmo_bilinear.restype = C.c_double
        L.dmo_bilinear.argtypes = [_vp, C.c_size_t, C.c_double, C.c_double]
        L.dmo_ncc.restype = C.c_double
        L.dmo_ncc.argtypes = [_vp, C.c_size_t, _vp, C.c_size_t, C.c_double, C.c_double, C.c_double, C.c_double]
        L.dmo_epipolar_search.argtypes = [_P(Params), _vp, C.c_size_t, _vp, C.c_size_t, _P(C.c_double), _P(C.c_double),
                                          C.c_double, C.c_double, C.c_double, C.c_double, _P(C.c_double)]
        L.dmo_update_depth_filter.argtypes = [_P(Params), _P(C.c_double), _P(C.c_double)] + [C.c_double] * 8 + [_P(C.c_double)]
        L.dmo_qr_solve2.restype = None
        L.dmo_qr_solve2.argtypes = [_P(C.c_double), _P(C.c_double), _P(C.c_double)]
        L.dmo_compose_T_C_R.restype = None
        L.dmo_compose_T_C_R.argtypes = [_P(C.c_double)] * 6
        L.dmo_transform_point.restype = None
        L.dmo_transform_point.argtypes = [_P(C.c_double)] * 4
        L.dmo_evaluate_depth.argtypes = [_P(Params), _vp, C.c_size_t, _vp, C.c_size_t, _vp, C.c_size_t, C.c_double,
                                         C.c_int, C.c_int, _P(C.c_double), _P(C.c_uint64)]
        L.dmo_variance_mask.argtypes = [C.c_int, C.c_int, _vp, C.c_size_t, C.c_double, _vp, C.c_size_t]
        L.dmo_point_cloud.restype = C.c_longlong
        L.dmo_point_cloud.argtypes = [_P(Params), _vp, C.c_size_t, C.c_int, _vp, C.c_size_t, _vp, C.c_size_t, _vp, _vp, C.c_longlong]
        L.dmo_max_threads.restype = C.c_int
        L.dmo_set_threads.restype = None
        L.dmo_set_threads.argtypes = [C.c_int]
        _libs["o"] = L
    return _libs["o"]


def ref_lib() -> Optional[C.CDLL]:
    """The compiled reference TU (640x480 only), or None where it was never built."""
    if "r" not in _libs:
        path = HERE / "_ref" / "libdmf_ref.so"
        if not path.exists():
            _libs["r"] = None
        else:
            L = C.CDLL(str(path))
            L.ref_update.restype = None
            L.ref_update.argtypes = [_vp, C.c_size_t, _vp, C.c_size_t, _P(C.c_double), _P(C.c_double), _vp, C.c_size_t, _vp, C.c_size_t]
            L.ref_ncc.restype = C.c_double
            L.ref_ncc.argtypes = [_vp, C.c_size_t, _vp, C.c_size_t, C.c_double, C.c_double, C.c_double, C.c_double]
            L.ref_epipolar_search.argtypes = [_vp, C.c_size_t, _vp, C.c_size_t, _P(C.c_double), _P(C.c_double),
                                              C.c_double, C.c_double, C.c_double, C.c_double, _P(C.c_double)]
            L.ref_update_depth_filter.argtypes = [_P(C.c_double), _P(C.c_double), C.c_int, C.c_int] + [C.c_double] * 6 + [_P(C.c_double)]
            _libs["r"] = L
    return _libs["r"]


def _d4(v):
    return (C.c_double * 4)(*[float(x) for x in v])


def _d3(v):
    return (C.c_double * 3)(*[float(x) for x in v])


def to_params(p) -> Params:
    """Accepts oracle.Params or slamplay_b200's DmfParams (same layout)."""
    out = Params()
    C.memmove(C.byref(out), C.byref(p), C.sizeof(Params))
    return out


def default_params(width=640, height=480, inverse_depth=False) -> Params:
    p = Params()
    assert lib().dmo_default_params(C.byref(p), width, height, int(inverse_depth)) == 0
    return p


def update(params, ref: np.ndarray, curr: np.ndarray, q, t, depth: np.ndarray, cov2: np.ndarray, *,
           rows=None, row_stride: int = 1, heap: bool = False, counters: Optional[Counters] = None,
           flags: Optional[np.ndarray] = None, dbg_ncc: Optional[np.ndarray] = None,
           dbg_n: Optional[np.ndarray] = None) -> None:
    """One reference update() (ref:355-393) in place on depth/cov2 (float64 H x W)."""
    p = to_params(params)
    r0, r1 = rows if rows is not None else (0, p.height)
    assert ref.dtype == np.uint8 and curr.dtype == np.uint8 and depth.dtype == np.float64 and cov2.dtype == np.float64
    rc = lib().dmo_update(C.byref(p), ref.ctypes.data, ref.strides[0], curr.ctypes.data, curr.strides[0], _d4(q), _d3(t),
                          depth.ctypes.data, depth.strides[0], cov2.ctypes.data, cov2.strides[0], r0, r1, row_stride,
                          int(heap), C.byref(counters) if counters is not None else None,
                          flags.ctypes.data if flags is not None else None, flags.strides[0] if flags is not None else 0,
                          dbg_ncc.ctypes.data if dbg_ncc is not None else None,
                          dbg_n.ctypes.data if dbg_n is not None else None)
    if rc != 0:
        raise RuntimeError("dmo_update failed")


def update_ex(params, ref, curr, q, t, depth, cov2, *, rows=None, row_stride=1, counters=None, flags=None,
              dbg_k=None, dbg_ncc64=None) -> None:
    """update() plus the diagnostic planes of dmo_update_ex (trip count << 16 | winning iteration; best NCC as f64)."""
    p = to_params(params)
    r0, r1 = rows if rows is not None else (0, p.height)
    rc = lib().dmo_update_ex(C.byref(p), ref.ctypes.data, ref.strides[0], curr.ctypes.data, curr.strides[0], _d4(q), _d3(t),
                             depth.ctypes.data, depth.strides[0], cov2.ctypes.data, cov2.strides[0], r0, r1, row_stride,
                             C.byref(counters) if counters is not None else None,
                             flags.ctypes.data if flags is not None else None, flags.strides[0] if flags is not None else 0,
                             dbg_k.ctypes.data if dbg_k is not None else None,
                             dbg_ncc64.ctypes.data if dbg_ncc64 is not None else None)
    if rc != 0:
        raise RuntimeError("dmo_update_ex failed")


def ref_update(ref: np.ndarray, curr: np.ndarray, q, t, depth: np.ndarray, cov2: np.ndarray) -> None:
    L = ref_lib()
    if L is None:
        raise RuntimeError("oracle/_ref/libdmf_ref.so was not built")
    assert ref.shape == (480, 640) and curr.shape == (480, 640)
    L.ref_update(ref.ctypes.data, ref.strides[0], curr.ctypes.data, curr.strides[0], _d4(q), _d3(t),
                 depth.ctypes.data, depth.strides[0], cov2.ctypes.data, cov2.strides[0])
