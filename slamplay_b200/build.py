"""In-tree build of the native libraries (no JIT cache: the .so files travel with the repo).

    python -m slamplay_b200.build          # libdmf.so (CUDA, sm_100a) + libdmf_synth_cpu.so

nvcc cross-compiles for sm_100a without a GPU.  `-lineinfo` keeps the ncu source page usable.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
ROOT = PKG.parent

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC", "-Xptxas", "-v",
]


def _host_cxx() -> str:
    # The image exports CXX=/opt/gcc/bin/g++ (a wrapper without libgomp.spec); prefer the system g++.
    return "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else (shutil.which("g++") or "g++")


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", shutil.which("nvcc")):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: the CUDA extension cannot be built (there is no CPU fallback)")


def _run(cmd: list[str], log: Path | None = None) -> str:
    proc = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if log is not None:
        log.write_text(" ".join(cmd) + "\n" + proc.stdout)
    if proc.returncode != 0:
        sys.stderr.write(proc.stdout)
        raise RuntimeError(f"build step failed ({proc.returncode}): {' '.join(cmd)}")
    return proc.stdout


def _stale(target: Path, sources: list[Path]) -> bool:
    if not target.exists():
        return True
    t = target.stat().st_mtime
    return any(s.stat().st_mtime > t for s in sources)


def build_cuda(force: bool = False, verbose: bool = False) -> Path:
    """libdmf.so = C ABI (include/dmf.h): kernels, frame ring, pipe micro-benchmarks — the product library."""
    target = PKG / "libdmf.so"
    srcs = [CSRC / "dmf_api.cu", CSRC / "dmf_kernels.cuh", CSRC / "dmf_geometry.h", CSRC / "microbench.cu", CSRC / "frame_ring.cu",
            CSRC / "dmf_internal.h", ROOT / "include" / "dmf.h"]
    if not force and not _stale(target, srcs):
        return target
    nvcc = _nvcc()
    build = PKG / "build"
    build.mkdir(exist_ok=True)
    ccbin = ["-ccbin", _host_cxx()]
    extra = os.environ.get("DMF_NVCC_EXTRA", "").split()  # tuning experiments, e.g. -DDMF_NCC_MIN_BLOCKS=3
    # -fopenmp: the host-side row copies / compares of dmf_update_strict (16*W*H bytes per call) run on a few threads
    out1 = _run([nvcc, *NVCC_FLAGS, "-Xcompiler", "-fopenmp", *extra, *ccbin, "-c", str(CSRC / "dmf_api.cu"), "-o", str(build / "dmf_api.o")],
                build / "dmf_api.ptxas.log")
    _run([nvcc, *NVCC_FLAGS, *ccbin, "-c", str(CSRC / "microbench.cu"), "-o", str(build / "microbench.o")], build / "microbench.ptxas.log")
    _run([nvcc, *NVCC_FLAGS, *ccbin, "-c", str(CSRC / "frame_ring.cu"), "-o", str(build / "frame_ring.o")], build / "frame_ring.ptxas.log")
    _run([nvcc, "-shared", *ccbin, "-o", str(target), str(build / "dmf_api.o"), str(build / "microbench.o"),
          str(build / "frame_ring.o"), "-lcudart", "-lrt", "-lgomp"])
    if verbose:
        print(out1)
    return target


def build_synth_cuda(force: bool = False) -> Path:
    """libdmf_synth.so = the CUDA renderer of the synthetic input sequences (include/dmf_synth.h).  An INPUT GENERATOR
    for tests and benchmarks, kept out of the product library so that a process which only needs inputs (bench.py
    --impl reference) never loads libdmf.so."""
    target = PKG / "libdmf_synth.so"
    srcs = [CSRC / "synth.cu", CSRC / "synth_scene.h", ROOT / "include" / "dmf_synth.h"]
    if not force and not _stale(target, srcs):
        return target
    nvcc = _nvcc()
    build = PKG / "build"
    build.mkdir(exist_ok=True)
    ccbin = ["-ccbin", _host_cxx()]
    # the renderer must not contract a*b+c into FMA (bit-identical with the g++ build)
    _run([nvcc, *NVCC_FLAGS, *ccbin, "-fmad=false", "-c", str(CSRC / "synth.cu"), "-o", str(build / "synth.o")], build / "synth.ptxas.log")
    _run([nvcc, "-shared", *ccbin, "-o", str(target), str(build / "synth.o"), "-lcudart"])
    return target


def build_synth_cpu(force: bool = False) -> Path:
    target = PKG / "libdmf_synth_cpu.so"
    srcs = [CSRC / "synth_cpu.cpp", CSRC / "synth_scene.h", ROOT / "include" / "dmf_synth.h"]
    if not force and not _stale(target, srcs):
        return target
    _run([_host_cxx(), "-O2", "-march=x86-64-v3", "-ffp-contract=off", "-fopenmp", "-fPIC", "-shared", "-std=c++17",
          "-o", str(target), str(CSRC / "synth_cpu.cpp")])
    return target


def build_all(force: bool = False, verbose: bool = False) -> dict[str, Path]:
    return {"libdmf": build_cuda(force, verbose), "libdmf_synth": build_synth_cuda(force), "libdmf_synth_cpu": build_synth_cpu(force)}


if __name__ == "__main__":
    res = build_all(force="--force" in sys.argv, verbose=True)
    for k, v in res.items():
        print(f"{k}: {v}")
