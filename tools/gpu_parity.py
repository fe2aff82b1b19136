"""GPU-vs-oracle parity run on a synthetic sequence (development tool; the judged tests live in tests/).

    python tools/gpu_parity.py [workload] [n_frames] [row_stride]
"""
import sys, time, json
import numpy as np
sys.path.insert(0, ".")
import oracle
from slamplay_b200.synth import make_sequence
from slamplay_b200.depth_filter import DepthFilter


def parity_metrics(p, d_gpu, c_gpu, d_ref, c_ref, rows=None):
    b = p.border
    ys = np.arange(b, p.height - b) if rows is None else np.array([y for y in rows if b <= y < p.height - b])
    dg, cg, dr, cr = (a[ys][:, b:p.width - b] for a in (d_gpu, c_gpu, d_ref, c_ref))
    both_nan = np.isnan(dg) & np.isnan(dr)
    ok = (np.abs(dg - dr) <= 1e-3 * np.abs(dr)) | both_nan
    conv_ref = cr < p.min_cov
    cls = lambda c: np.where(np.isnan(c), 3, np.where(c < p.min_cov, 0, np.where(c > p.max_cov, 1, 2)))
    return {
        "P1_depth_within_1e-3": float(ok.mean()),
        "P1_converged_only": float(ok[conv_ref].mean()) if conv_ref.any() else None,
        "P2_final_class_mismatch": float((cls(cg) != cls(cr)).mean()),
        "exact_equal_frac": float(((dg == dr) | both_nan).mean()),
        "within_1e-6": float(((np.abs(dg - dr) <= 1e-6 * np.abs(dr)) | both_nan).mean()),
        "max_rel_cov": float(np.nanmax(np.abs(cg - cr) / np.maximum(np.abs(cr), 1e-300))),
        "converged_frac_ref": float(conv_ref.mean()),
    }


def main():
    wl = sys.argv[1] if len(sys.argv) > 1 else "remode_640x480"
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 20
    stride = int(sys.argv[3]) if len(sys.argv) > 3 else 1
    seq = make_sequence(wl, n_frames=n)
    p = seq.params
    h, w = seq.shape
    t0 = time.time()
    frames = [seq.render_host(i) for i in range(n)]
    print(f"rendered {n} frames {w}x{h} on CPU in {time.time()-t0:.1f}s", flush=True)
    f = DepthFilter(p)
    f.set_reference(frames[0])
    f.fill_state(3.0, 3.0)
    f.enable_flags(True)
    d_ref = np.full((h, w), 3.0); c_ref = np.full((h, w), 3.0)
    rows = range(p.border, h - p.border, stride)
    fl_ref = np.zeros((h, w), np.uint8)
    oc = oracle.Counters()
    worst = 0.0
    for i in range(1, n):
        T = seq.T_C_R(i)
        f.update(frames[i], T)
        fl_gpu = f.flags()
        oracle.update(p, frames[0], frames[i], T.q, T.t, d_ref, c_ref, rows=(p.border, h - p.border), row_stride=stride,
                      counters=oc, flags=fl_ref)
        ys = np.array(list(rows))
        mism = float((fl_gpu[ys][:, p.border:w - p.border] != fl_ref[ys][:, p.border:w - p.border]).mean())
        worst = max(worst, mism)
        if i <= 5 or i % 10 == 0:
            d_gpu, c_gpu = f.download_state()
            m = parity_metrics(p, d_gpu, c_gpu, d_ref, c_ref, rows)
            print(i, "flag mismatch %.2e" % mism, json.dumps(m), flush=True)
    d_gpu, c_gpu = f.download_state()
    m = parity_metrics(p, d_gpu, c_gpu, d_ref, c_ref, rows)
    print("FINAL", json.dumps(m), "worst per-frame decision mismatch %.3e" % worst)
    print("gpu counters", f.counters(), "oracle counters", oc.as_dict())


if __name__ == "__main__":
    main()
