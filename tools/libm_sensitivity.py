"""How far can a libm that is not bit-identical to glibc move the depth filter?  (development tool, CPU only)

The CUDA path runs the reference's arithmetic operation by operation (slamplay_b200/csrc/dmf_geometry.h, bit-identical to
the oracle on the host) except for acos / sin at ref:527-533, where CUDA's libm and glibc may differ by 1-2 ulp.  This
tool runs the ORACLE twice on the same sequence — once as is, once with the two acos results moved by up to +-N ulp
pseudo-randomly — and reports the drift between the two runs in the terms of tools/parity_diag.py, so that the residual
GPU-vs-oracle differences can be compared with what libm rounding alone produces.

    python tools/libm_sensitivity.py WORKLOAD [--frames F] [--rows N] [--ulp 1]
"""
import argparse
import json
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import oracle  # noqa: E402
from slamplay_b200.synth import make_sequence  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("workload")
ap.add_argument("--frames", type=int, default=None)
ap.add_argument("--rows", type=int, default=16)
ap.add_argument("--ulp", type=int, default=1)
ap.add_argument("--out", default=None)
a = ap.parse_args()
seq = make_sequence(a.workload, n_frames=a.frames)
p = seq.params
h, w = seq.shape
b = p.border
stride = max(1, (h - 2 * b) // a.rows)
r0 = b + stride // 2
ys = np.arange(r0, h - b, stride)[: a.rows]
frames = [seq.render_host(i) for i in range(seq.n_frames)]
L = oracle.lib()
L.dmo_set_libm_perturbation.argtypes = [__import__("ctypes").c_int]
state = [[np.full((h, w), 3.0), np.full((h, w), 3.0)] for _ in range(2)]
flag_diff_pixels = np.zeros((len(ys), w - 2 * b), bool)
hist = []
for i in range(1, seq.n_frames):
    T = seq.T_C_R(i)
    fl = []
    for k, ulp in enumerate((0, a.ulp)):
        L.dmo_set_libm_perturbation(ulp)
        f = np.zeros((h, w), np.uint8)
        oracle.update(p, frames[0], frames[i], T.q, T.t, state[k][0], state[k][1], rows=(int(ys[0]), int(ys[-1]) + 1), row_stride=stride, flags=f)
        fl.append(f[ys][:, b:w - b])
    L.dmo_set_libm_perturbation(0)
    flag_diff_pixels |= fl[0] != fl[1]
    d0, d1 = state[0][0][ys][:, b:w - b], state[1][0][ys][:, b:w - b]
    rel = np.abs(d1 - d0) / np.maximum(np.abs(d0), 1e-300)
    rel = np.where(np.isnan(d0) & np.isnan(d1), 0.0, np.where(np.isnan(rel), np.inf, rel))
    hist.append({"update": i, "frac_rel>1e-12": float((rel > 1e-12).mean()), "frac_rel>1e-9": float((rel > 1e-9).mean()),
                 "frac_rel>1e-6": float((rel > 1e-6).mean()), "frac_rel>1e-3": float((rel > 1e-3).mean()), "median_rel": float(np.median(rel))})
    if i <= 3 or i % 20 == 0 or i == seq.n_frames - 1:
        print(hist[-1], flush=True)
out = {"workload": a.workload, "updates": seq.n_frames - 1, "rows": int(len(ys)), "perturbation_ulp": a.ulp,
       "pixels": int(flag_diff_pixels.size), "pixels_with_a_decision_difference": int(flag_diff_pixels.sum()),
       "final": hist[-1], "history": hist[:: max(1, len(hist) // 25)]}
print(json.dumps(out, indent=1))
if a.out:
    Path(a.out).write_text(json.dumps(out, indent=1))
