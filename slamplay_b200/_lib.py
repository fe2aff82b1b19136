"""ctypes binding of the C ABI declared in include/dmf.h / include/dmf_synth.h.

The CUDA library is REQUIRED: importing symbols from a missing / unbuildable libdmf.so raises.
There is no CPU or PyTorch fallback for the depth-filter path.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

PKG = Path(__file__).resolve().parent


class DmfParams(C.Structure):
    """struct dmf_params (include/dmf.h) — runtime form of the reference constants
    dense_mapping/test_monocular_mapping.cpp:72-89."""

    _fields_ = [
        ("width", C.c_int32), ("height", C.c_int32), ("border", C.c_int32), ("ncc_half", C.c_int32),
        ("fx", C.c_double), ("fy", C.c_double), ("cx", C.c_double), ("cy", C.c_double),
        ("step", C.c_double), ("max_half_len", C.c_double), ("min_depth", C.c_double), ("n_sigma", C.c_double),
        ("ncc_thresh", C.c_double), ("min_cov", C.c_double), ("max_cov", C.c_double),
        ("inverse_depth", C.c_int32), ("reserved", C.c_int32),
    ]

    def copy(self) -> "DmfParams":
        out = DmfParams()
        C.memmove(C.byref(out), C.byref(self), C.sizeof(DmfParams))
        return out


class DmfCounters(C.Structure):
    _fields_ = [("frames", C.c_uint64), ("interior", C.c_uint64), ("active", C.c_uint64),
                ("ncc_evals", C.c_uint64), ("accepted", C.c_uint64)]

    def as_dict(self) -> dict:
        return {k: int(getattr(self, k)) for k, _ in self._fields_}


class DmfPipeRate(C.Structure):
    _fields_ = [("per_clk_sm", C.c_double), ("per_second", C.c_double), ("eff_mhz", C.c_double)]


class DmfPipePeaks(C.Structure):
    _fields_ = [("n_sm", C.c_int32), ("reserved", C.c_int32), ("ffma", DmfPipeRate), ("dfma", DmfPipeRate),
                ("idp4a", DmfPipeRate), ("i2f_f64", DmfPipeRate), ("ldg64_l1", DmfPipeRate), ("ldg128_l1", DmfPipeRate)]

    def as_dict(self) -> dict:
        out = {"n_sm": int(self.n_sm)}
        for k in ("ffma", "dfma", "idp4a", "i2f_f64", "ldg64_l1", "ldg128_l1"):
            r = getattr(self, k)
            out[k] = {"per_clk_sm": r.per_clk_sm, "per_second": r.per_second, "eff_mhz": r.eff_mhz}
        return out


class SynthScene(C.Structure):
    _fields_ = [("plane_z", C.c_double), ("relief_amp", C.c_double), ("relief_period", C.c_double),
                ("tex_base", C.c_double), ("tex_octaves", C.c_int32), ("seed", C.c_uint32),
                ("ray_iters", C.c_int32), ("supersample", C.c_int32)]


class SynthCamera(C.Structure):
    _fields_ = [("width", C.c_int32), ("height", C.c_int32), ("fx", C.c_double), ("fy", C.c_double),
                ("cx", C.c_double), ("cy", C.c_double), ("q", C.c_double * 4), ("t", C.c_double * 3)]


# every symbol include/dmf.h declares, with its ctypes signature
_P = C.POINTER
_vp = C.c_void_p
DMF_SYMBOLS = {
    "dmf_abi_version": (C.c_int, []),
    "dmf_build_info": (C.c_char_p, []),
    "dmf_last_error": (C.c_char_p, [_vp]),
    "dmf_default_params": (C.c_int, [_P(DmfParams), C.c_int, C.c_int, C.c_int]),
    "dmf_create": (C.c_int, [_P(DmfParams), C.c_int, C.c_int, C.c_int, _P(_vp)]),
    "dmf_create_cyclic": (C.c_int, [_P(DmfParams), C.c_int, C.c_int, C.c_int, C.c_int, _P(_vp)]),
    "dmf_get_rows": (C.c_int, [_vp, _P(C.c_int), C.c_int, _P(C.c_int)]),
    "dmf_destroy": (None, [_vp]),
    "dmf_get_params": (C.c_int, [_vp, _P(DmfParams)]),
    "dmf_get_band": (C.c_int, [_vp, _P(C.c_int), _P(C.c_int)]),
    "dmf_set_reference": (C.c_int, [_vp, _vp, C.c_size_t]),
    "dmf_set_reference_device": (C.c_int, [_vp, _vp, C.c_size_t]),
    "dmf_fill_state": (C.c_int, [_vp, C.c_double, C.c_double]),
    "dmf_upload_state": (C.c_int, [_vp, _vp, C.c_size_t, _vp, C.c_size_t]),
    "dmf_download_state": (C.c_int, [_vp, _vp, C.c_size_t, _vp, C.c_size_t]),
    "dmf_update": (C.c_int, [_vp, _vp, C.c_size_t, _P(C.c_double), _P(C.c_double)]),
    "dmf_update_device": (C.c_int, [_vp, _vp, C.c_size_t, _P(C.c_double), _P(C.c_double), _vp]),
    "dmf_update_strict": (C.c_int, [_vp, _vp, C.c_size_t, _vp, C.c_size_t, _P(C.c_double), _P(C.c_double), _vp, C.c_size_t, _vp, C.c_size_t]),
    "dmf_flush": (C.c_int, [_vp]),
    "dmf_sync": (C.c_int, [_vp]),
    "dmf_set_timing": (C.c_int, [_vp, C.c_int]),
    "dmf_get_timing": (C.c_int, [_vp, _P(C.c_double), _P(C.c_uint64), C.c_int]),
    "dmf_read_counters": (C.c_int, [_vp, _P(DmfCounters), C.c_int]),
    "dmf_enable_flags": (C.c_int, [_vp, C.c_int]),
    "dmf_download_flags": (C.c_int, [_vp, _vp, C.c_size_t]),
    "dmf_download_debug": (C.c_int, [_vp, _vp, _vp]),
    "dmf_device_state": (C.c_int, [_vp, _P(_vp), _P(_vp), _P(C.c_size_t)]),
    "dmf_stream": (C.c_int, [_vp, _P(_vp)]),
    "dmf_selftest_division": (C.c_int, [C.c_int, C.c_uint64, C.c_uint64, _P(C.c_uint64)]),
    "dmf_alloc_pinned": (C.c_int, [_P(_vp), C.c_size_t]),
    "dmf_free_pinned": (C.c_int, [_vp]),
    "dmf_host_register": (C.c_int, [_vp, C.c_size_t]),
    "dmf_host_unregister": (C.c_int, [_vp]),
    "dmf_set_truth": (C.c_int, [_vp, _vp, C.c_size_t]),
    "dmf_evaluate_depth": (C.c_int, [_vp, C.c_double, _P(C.c_double), _P(C.c_uint64)]),
    "dmf_variance_mask": (C.c_int, [_vp, C.c_double, _vp, C.c_size_t]),
    "dmf_ring_create": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _P(_vp), _vp]),
    "dmf_ring_open": (C.c_int, [C.c_int, _vp, C.c_int, _P(_vp)]),
    "dmf_ring_close": (None, [_vp]),
    "dmf_ring_info": (C.c_int, [_vp, _P(C.c_int), _P(C.c_int), _P(C.c_uint32)]),
    "dmf_ring_publish": (C.c_int, [_vp, _vp, C.c_size_t, _vp]),
    "dmf_update_ring": (C.c_int, [_vp, _vp, _P(C.c_double), _P(C.c_double)]),
    "dmf_pipe_peaks": (C.c_int, [C.c_int, _P(DmfPipePeaks)]),
    "dmf_point_cloud": (C.c_int, [_vp, _vp, C.c_size_t, C.c_int, C.c_double, _vp, _vp, C.c_uint64, _P(C.c_uint64)]),
}
SYNTH_DEVICE_SYMBOLS = {
    "dmf_synth_render_device": (C.c_int, [_P(SynthScene), _P(SynthCamera), _vp, C.c_size_t, _vp, C.c_size_t, _vp]),
}
SYNTH_HOST_SYMBOLS = {
    "dmf_synth_render_host": (C.c_int, [_P(SynthScene), _P(SynthCamera), _vp, C.c_size_t, _vp, C.c_size_t]),
}

_cache: dict[str, C.CDLL] = {}


def _bind(lib: C.CDLL, table: dict) -> None:
    for name, (res, args) in table.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is missing: fail loudly
        fn.restype = res
        fn.argtypes = args


def load_dmf() -> C.CDLL:
    """Load slamplay_b200/libdmf.so (CUDA, sm_100a).  Raises if it is not built."""
    if "dmf" not in _cache:
        # DMF_LIB: development override (A/B runs of kernel variants built by tools/build_variants.sh)
        path = Path(os.environ["DMF_LIB"]) if os.environ.get("DMF_LIB") else PKG / "libdmf.so"
        if not path.exists():
            raise RuntimeError(
                f"{path} is missing: build it with `python -m slamplay_b200.build` "
                "(needs nvcc). The depth-filter path has no CPU fallback.")
        lib = C.CDLL(str(path))
        _bind(lib, DMF_SYMBOLS)
        _cache["dmf"] = lib
    return _cache["dmf"]


def load_synth_cuda() -> C.CDLL:
    """slamplay_b200/libdmf_synth.so: the CUDA renderer of the synthetic inputs (not part of the product library)."""
    if "synth_cuda" not in _cache:
        path = PKG / "libdmf_synth.so"
        if not path.exists():
            raise RuntimeError(f"{path} is missing: build it with `python -m slamplay_b200.build`")
        lib = C.CDLL(str(path))
        _bind(lib, SYNTH_DEVICE_SYMBOLS)
        _cache["synth_cuda"] = lib
    return _cache["synth_cuda"]


def load_synth_cpu() -> C.CDLL:
    if "synth_cpu" not in _cache:
        path = PKG / "libdmf_synth_cpu.so"
        if not path.exists():
            raise RuntimeError(f"{path} is missing: build it with `python -m slamplay_b200.build`")
        lib = C.CDLL(str(path))
        _bind(lib, SYNTH_HOST_SYMBOLS)
        _cache["synth_cpu"] = lib
    return _cache["synth_cpu"]
