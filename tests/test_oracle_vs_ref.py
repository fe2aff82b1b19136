"""Oracle restatement vs the compiled reference translation unit, live (only where oracle/_ref was
built, i.e. in the container that has /root/reference; the frozen form of this check is
tests/test_oracle_golden.py)."""
import numpy as np
import pytest

import oracle


@pytest.fixture(scope="module")
def ref():
    L = oracle.ref_lib()
    if L is None:
        pytest.skip("oracle/_ref/libdmf_ref.so not built (no /root/reference here)")
    return L


def test_update_bit_equal_on_random_state(ref, seq640):
    """Random but plausible state maps (incl. converged, diverged and NaN pixels): one update() each."""
    seq, frames = seq640
    h, w = seq.shape
    rng = np.random.default_rng(7)
    for i in (1, 3, 5):
        depth = rng.uniform(1.5, 3.5, (h, w))
        cov2 = 10.0 ** rng.uniform(-5, 1.2, (h, w))  # spans min_cov=1e-4 .. max_cov=10
        depth[::37, ::41] = np.nan
        cov2[::53, ::29] = np.nan
        d2, c2 = depth.copy(), cov2.copy()
        T = seq.T_C_R(i)
        oracle.update(seq.params, frames[0], frames[i], T.q, T.t, depth, cov2)
        oracle.ref_update(frames[0], frames[i], T.q, T.t, d2, c2)
        assert np.array_equal(depth, d2, equal_nan=True)
        assert np.array_equal(cov2, c2, equal_nan=True)


def test_degenerate_poses_bit_equal(ref, seq640):
    """Zero baseline (acos(0/0) NaN poison, ref:527) and pure rotation."""
    seq, frames = seq640
    h, w = seq.shape
    for q, t in [((0, 0, 0, 1), (0, 0, 0)), ((0.01, -0.02, 0.005, 0.9997), (0, 0, 0)), ((0, 0, 0, 1), (1e-9, 0, 0))]:
        q = np.array(q, float)
        q /= np.linalg.norm(q)
        d1, c1 = np.full((h, w), 2.0), np.full((h, w), 0.5)
        d2, c2 = d1.copy(), c1.copy()
        oracle.update(seq.params, frames[0], frames[0], q, t, d1, c1, rows=(200, 216))
        oracle.ref_update(frames[0], frames[0], q, t, d2, c2)
        assert np.array_equal(d1[200:216], d2[200:216], equal_nan=True)
        assert np.array_equal(c1[200:216], c2[200:216], equal_nan=True)
