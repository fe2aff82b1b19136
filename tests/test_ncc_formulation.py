"""CPU restatement of the arithmetic the CUDA path uses for one NCC evaluation (moments_kernel + ncc_kernel,
slamplay_b200/csrc/dmf_kernels.cuh) checked against the oracle's two-pass FP64 ZNCC (reference NCC ref:449-480 with
the bilinear taps of ref:165-174).

All 49 taps of one evaluation share the same bilinear weights, so every sum is a linear / quadratic form in the
weights over INTEGER moments of the 8x8 block under the sample; the kernel keeps those exact in int32 and only the
final combination runs in FP64.  This test pins the algebra (and the int32 range claims) without a GPU."""
import numpy as np

import oracle

EPS_INT = 1015.2029750625  # 1e-10 * (49 * 255^2)^2: the reference's epsilon (ref:479) in centred-integer units
I32 = 2 ** 31


def block_moments(B):
    """Frame-only moments of the 8x8 block B (what moments_kernel stores for the block position): window sums S[a][b],
    centred squares q[a][b], centred horizontal / vertical neighbour products H[b], V[a], summed diagonal term D."""
    B = B.astype(np.int64)
    W = {(a, b): B[b:b + 7, a:a + 7] for a in (0, 1) for b in (0, 1)}  # window (a, b): columns a..a+6, rows b..b+6
    S = {k: int(w.sum()) for k, w in W.items()}
    q = {k: 49 * int((w * w).sum()) - S[k] ** 2 for k, w in W.items()}
    H = [49 * int((W[0, b] * W[1, b]).sum()) - S[0, b] * S[1, b] for b in (0, 1)]
    V = [49 * int((W[a, 0] * W[a, 1]).sum()) - S[a, 0] * S[a, 1] for a in (0, 1)]
    D1 = 49 * int((W[0, 0] * W[1, 1]).sum()) - S[0, 0] * S[1, 1]
    D2 = 49 * int((W[1, 0] * W[0, 1]).sum()) - S[1, 0] * S[0, 1]
    return W, S, q, H, V, D1, D2


def ncc_from_moments(patch, B, fx, fy):
    """ncc_combine(): separable form of num = w.cR and den2 = w^T G w with w = (gx gy, fx gy, gx fy, fx fy)."""
    r = patch.astype(np.int64)
    Sr = int(r.sum())
    den1 = 49 * int((r * r).sum()) - Sr * Sr
    W, S, q, H, V, D1, D2 = block_moments(B)
    cR = {k: 49 * int((r * w).sum()) - Sr * S[k] for k, w in W.items()}
    ints = [den1, D1, D2, D1 + D2, *q.values(), *H, *V, *cR.values()]
    assert all(-I32 <= v < I32 for v in ints), "a centred moment leaves int32"
    gx, gy = 1.0 - fx, 1.0 - fy
    n0 = gx * cR[0, 0] + fx * cR[1, 0]
    n1 = gx * cR[0, 1] + fx * cR[1, 1]
    num = fy * n1 + gy * n0
    A, Bm, Cc, Dd, E, F = gx * gx, gx * fx, fx * fx, gy * gy, gy * fy, fy * fy
    t0 = A * q[0, 0] + Cc * q[1, 0] + 2 * Bm * H[0]
    t1 = A * q[0, 1] + Cc * q[1, 1] + 2 * Bm * H[1]
    t2 = A * V[0] + Cc * V[1] + Bm * (D1 + D2)
    den2 = Dd * t0 + F * t1 + 2 * E * t2
    return num / np.sqrt(den1 * den2 + EPS_INT)


def _oracle_ncc(ref, curr, rx, ry, cx, cy):
    return oracle.lib().dmo_ncc(ref.ctypes.data, ref.strides[0], curr.ctypes.data, curr.strides[0],
                                float(rx), float(ry), float(cx), float(cy))


def _check(ref, curr, rx, ry, cx, cy, tol=2e-12):
    ix, iy = int(np.floor(cx)), int(np.floor(cy))
    patch = ref[ry - 3:ry + 4, rx - 3:rx + 4]
    B = curr[iy - 3:iy + 5, ix - 3:ix + 5]
    got = ncc_from_moments(patch, B, cx - ix, cy - iy)
    want = _oracle_ncc(ref, curr, rx, ry, cx, cy)
    assert abs(got - want) <= tol, (got, want, cx, cy)
    return got


def test_integer_moment_form_matches_the_two_pass_zncc_on_random_texture():
    rng = np.random.default_rng(7)
    ref = rng.integers(0, 256, (64, 96), dtype=np.uint8)
    curr = np.clip(ref.astype(np.int32) + rng.integers(-30, 31, ref.shape), 0, 255).astype(np.uint8)
    for _ in range(400):
        rx, ry = int(rng.integers(8, 88)), int(rng.integers(8, 56))
        cx, cy = rng.uniform(8, 87), rng.uniform(8, 55)
        _check(ref, curr, rx, ry, cx, cy)


def test_integer_positions_and_fractions_at_the_ends():
    rng = np.random.default_rng(8)
    ref = rng.integers(0, 256, (48, 48), dtype=np.uint8)
    curr = rng.integers(0, 256, (48, 48), dtype=np.uint8)
    for cx, cy in [(20.0, 20.0), (20.0, 21.5), (21.5, 20.0), (20.0 + 1e-12, 20.0), (21.0 - 1e-12, 22.0 - 1e-12),
                   (20.7, 20.0), (20.999999999, 20.000000001)]:
        _check(ref, curr, 24, 24, cx, cy)
    assert abs(_check(ref, ref, 24, 24, 24.0, 24.0) - 1.0) < 1e-9  # the patch against itself


def test_extreme_blocks_stay_inside_int32_and_match():
    yy, xx = np.mgrid[0:48, 0:48]
    checker = (((xx + yy) & 1) * 255).astype(np.uint8)
    stripes = ((xx & 1) * 255).astype(np.uint8)
    white = np.full((48, 48), 255, np.uint8)
    black = np.zeros((48, 48), np.uint8)
    ramp = np.clip(xx * 6, 0, 255).astype(np.uint8)
    imgs = [checker, stripes, white, black, ramp]
    for ref in imgs:
        for curr in imgs:
            for cx, cy in [(24.0, 24.0), (24.5, 24.0), (24.5, 24.5), (24.25, 24.75)]:
                v = _check(ref, curr, 24, 24, cx, cy, tol=1e-9)
                assert -1.0 - 1e-9 <= v <= 1.0 + 1e-9
    # flat against anything: numerator 0, denominator sqrt(eps) -> exactly the reference's 0 / sqrt(1e-10) behaviour
    assert ncc_from_moments(white[21:28, 21:28], checker[21:29, 21:29], 0.5, 0.5) == 0.0
