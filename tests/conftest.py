import os
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


@pytest.fixture(scope="session", autouse=True)
def _native_libs():
    """Build the native pieces once (nvcc cross-compiles without a GPU).  The oracle is test
    infrastructure; oracle/_ref is only (re)built where the reference tree exists."""
    from slamplay_b200 import build as b

    b.build_all()
    import oracle

    oracle.build(ref=True)
    yield


@pytest.fixture(scope="session")
def has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.fixture(scope="session")
def seq640():
    """REMODE-shaped 640x480 synthetic sequence, 6 frames rendered on the CPU."""
    from slamplay_b200.synth import make_sequence

    seq = make_sequence("remode_640x480", n_frames=6)
    frames = [seq.render_host(i) for i in range(seq.n_frames)]
    return seq, frames
