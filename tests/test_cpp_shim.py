"""The C++ host shim (slamplay_b200/cpp/dense_mono_update.hpp) compiles against the C ABI; on a GPU the
example driver — the reference's loop ref:285-305 through update(ref, curr, T_C_R, depth, depth_cov2) —
runs and converges."""
import re
import subprocess
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
PKG = ROOT / "slamplay_b200"


def _build(tmp_path):
    exe = tmp_path / "example_sequence"
    cmd = ["/usr/bin/g++", "-std=c++17", "-O2", "-Wall", str(PKG / "cpp" / "example_sequence.cpp"), "-o", str(exe),
           f"-L{PKG}", "-ldmf", "-ldmf_synth_cpu", f"-Wl,-rpath,{PKG}", "-L/usr/local/cuda/lib64"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return exe


def test_shim_compiles_and_links(tmp_path):
    _build(tmp_path)


def test_shim_signature_is_the_reference_signature():
    txt = (PKG / "cpp" / "dense_mono_update.hpp").read_text()
    assert re.search(r"inline void update\(const Mat &ref, const Mat &curr, const SE3d &T_C_R, Mat &depth, Mat &depth_cov2\)", txt)


@pytest.mark.gpu
def test_example_driver_runs(tmp_path):
    exe = _build(tmp_path)
    r = subprocess.run([str(exe), "6"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    lines = [l for l in r.stdout.splitlines() if l.startswith("*** loop")]
    assert len(lines) == 5
    n_last = int(re.search(r"cov2 < 0.5: (\d+)", lines[-1]).group(1))
    mean = float(re.search(r"mean depth ([0-9.]+)", lines[-1]).group(1))
    assert n_last > 50000 and 1.7 < mean < 2.6
