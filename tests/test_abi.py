"""The C-ABI library loads and exports every symbol include/*.h declares; without a GPU every compute
entry point fails loudly (no CPU fallback); nothing in the product package touches oracle/."""
import ctypes as C
import re
from pathlib import Path

import pytest

from slamplay_b200 import _lib

ROOT = Path(__file__).resolve().parent.parent


def declared_symbols(header: Path):
    text = re.sub(r"/\*.*?\*/", "", header.read_text(), flags=re.S)
    return sorted(set(re.findall(r"\b(dmf_[a-z0-9_]+)\s*\(", text)))


def test_every_declared_symbol_is_exported_and_bound():
    lib = _lib.load_dmf()
    syms = declared_symbols(ROOT / "include" / "dmf.h")
    assert len(syms) >= 25
    for s in syms:
        assert hasattr(lib, s), f"libdmf.so does not export {s}"
        assert s in _lib.DMF_SYMBOLS, f"{s} has no ctypes signature in slamplay_b200/_lib.py"
    for s in declared_symbols(ROOT / "include" / "dmf_synth.h"):
        owner = _lib.load_synth_cpu() if s.endswith("_host") else _lib.load_synth_cuda()
        assert hasattr(owner, s), f"{s} is not exported"


def test_abi_version_and_build_info():
    lib = _lib.load_dmf()
    assert lib.dmf_abi_version() == 1
    assert b"sm_100a" in lib.dmf_build_info()


def test_params_struct_layout_matches_header():
    assert C.sizeof(_lib.DmfParams) == 4 * 4 + 11 * 8 + 2 * 4
    assert C.sizeof(_lib.DmfCounters) == 5 * 8


def test_default_params_reference_constants():
    lib = _lib.load_dmf()
    p = _lib.DmfParams()
    assert lib.dmf_default_params(C.byref(p), 640, 480, 0) == 0
    import oracle
    assert bytes(p) == bytes(oracle.default_params(640, 480))
    assert lib.dmf_default_params(C.byref(p), 1920, 1080, 1) == 0
    assert bytes(p) == bytes(oracle.default_params(1920, 1080, True))
    assert lib.dmf_default_params(None, 640, 480, 0) < 0
    assert b"bad arguments" in lib.dmf_last_error(None)


def test_create_validates_params_before_touching_cuda():
    lib = _lib.load_dmf()
    ctx = C.c_void_p()
    p = _lib.DmfParams()
    lib.dmf_default_params(C.byref(p), 640, 480, 0)
    p.ncc_half = 2
    assert lib.dmf_create(C.byref(p), 0, 0, 480, C.byref(ctx)) == -1 and not ctx.value
    assert b"ncc_half" in lib.dmf_last_error(None)
    p.ncc_half, p.border = 3, 2
    assert lib.dmf_create(C.byref(p), 0, 0, 480, C.byref(ctx)) == -1
    p.border = 20
    assert lib.dmf_create(C.byref(p), 0, 10, 5, C.byref(ctx)) == -1  # row_begin > row_end


def test_no_cpu_fallback(has_gpu):
    """Without a CUDA device the product path must fail loudly, not compute on the CPU."""
    if has_gpu:
        pytest.skip("GPU present")
    from slamplay_b200.depth_filter import DepthFilter, DmfError
    with pytest.raises(DmfError, match="no CUDA device|CPU fallback"):
        DepthFilter(width=640, height=480)


def test_product_package_never_references_the_oracle():
    for f in (ROOT / "slamplay_b200").rglob("*"):
        if f.suffix in (".py", ".cu", ".cuh", ".h", ".hpp", ".cpp") and f.is_file():
            txt = f.read_text()
            assert not re.search(r"^\s*(import|from)\s+oracle\b", txt, flags=re.M), f"{f} imports the oracle"
            assert "liboracle" not in txt and "dmo_" not in txt, f"{f} links against the oracle"


def test_ring_and_strict_entry_points_validate_arguments_without_a_gpu():
    """The multi-GPU frame ring and the strict drop-in reject bad arguments before touching CUDA; with valid arguments
    and no device they fail loudly (no CPU fallback)."""
    lib = _lib.load_dmf()
    ring = C.c_void_p()
    handle = (C.c_uint8 * 192)()
    assert lib.dmf_ring_create(0, 1, 640, 480, 2, C.byref(ring), handle) == -1      # n_slots < 2
    assert lib.dmf_ring_create(0, 4, 640, 480, 0, C.byref(ring), handle) == -1      # no consumers
    assert lib.dmf_ring_create(0, 4, 640, 480, 2, None, handle) == -1
    assert lib.dmf_ring_open(0, handle, 0, C.byref(ring)) == -1                     # zeroed bytes are not a ring handle
    assert b"not a ring handle" in lib.dmf_last_error(None)
    assert lib.dmf_update_ring(None, None, None, None) == -1
    assert lib.dmf_ring_publish(None, None, 0, None) == -1
    assert lib.dmf_update_strict(None, None, 0, None, 0, None, None, None, 0, None, 0) == -1
    assert lib.dmf_host_register(None, 0) == -1
    import torch
    if not torch.cuda.is_available():
        assert lib.dmf_ring_create(0, 4, 640, 480, 2, C.byref(ring), handle) == -2  # DMF_ERR_CUDA: no device, no fallback
        peaks = _lib.DmfPipePeaks()
        assert lib.dmf_pipe_peaks(0, C.byref(peaks)) == -2
        bad = C.c_uint64()
        assert lib.dmf_selftest_division(0, 10, 1, C.byref(bad)) == -2
