"""First-divergence diagnosis of GPU-vs-oracle deviations (development tool; VERDICT r01 "What's weak" 1).

For a sample of interior rows, the CUDA path (debug planes on: per-update set-up + fusion from the maps, bit-identical
to the deferred pipeline) and the CPU oracle run side by side over the whole sequence.  Per update and per sampled pixel
the tool compares: gate flag (ref:366), trip count of the l-loop (ref:432), index of the winning sample (ref:438-441),
accept flag (ref:443) and the fused state.  Every pixel whose FINAL depth deviates by more than 1e-3 (and every pixel
whose decisions ever differ) is attributed to the FIRST update at which anything observable differed, and classified:

  gate       the variance gate differs (cov2 crossed min_cov / max_cov one update apart)
  trip       same gate, different trip count (sample added / dropped at the segment end)
  argmax     same trip count, different winning sample (near-tie between two samples)
  accept     same winner, different NCC >= 0.85f decision
  drift      all decisions equal in every update, state drifted numerically

    python tools/parity_diag.py WORKLOAD [--frames F] [--rows N] [--out FILE] [--fake-gpu]
"""
import argparse
import json
import sys
import time
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import oracle  # noqa: E402
from slamplay_b200.synth import make_sequence  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("workload")
    ap.add_argument("--frames", type=int, default=None)
    ap.add_argument("--rows", type=int, default=64)
    ap.add_argument("--out", default=None)
    ap.add_argument("--fake-gpu", action="store_true", help="CPU syntax check: the oracle's heap variant plays the GPU")
    ap.add_argument("--examples", type=int, default=12)
    a = ap.parse_args()

    seq = make_sequence(a.workload, n_frames=a.frames)
    p = seq.params
    h, w = seq.shape
    F = seq.n_frames
    b = p.border
    lo, hi = b, h - b
    stride = max(1, (hi - lo) // a.rows)
    first = lo + stride // 2
    ys = np.arange(first, hi, stride)[: a.rows]
    xs = slice(b, w - b)
    t0 = time.time()
    if not a.fake_gpu:
        import torch
        from slamplay_b200.depth_filter import DepthFilter
        pitch = (w + 15) // 16 * 16
        dev = torch.zeros((F, h, pitch), dtype=torch.uint8, device="cuda")
        for i in range(F):
            seq.render_device(i, dev[i].data_ptr(), pitch, stream=torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        frames = dev[:, :, :w].cpu().numpy()
        f = DepthFilter(p)
        f.set_reference(frames[0])
        f.fill_state(3.0, 3.0)
        f.enable_flags(True)
    else:
        frames = [seq.render_host(i) for i in range(F)]
        d_g, c_g = np.full((h, w), 3.0), np.full((h, w), 3.0)
    print(f"[diag] {a.workload} {w}x{h} F={F}; {len(ys)} rows (every {stride}th from {first}); inputs in {time.time() - t0:.1f}s", flush=True)

    d_o, c_o = np.full((h, w), 3.0), np.full((h, w), 3.0)
    n_s = (len(ys), w - 2 * b)
    first_frame = np.full(n_s, -1, np.int32)      # first update with an observable decision difference
    first_class = np.zeros(n_s, np.int8)          # 1 gate 2 trip 3 argmax 4 accept
    first_info = {}                               # (row index, col) -> details
    pre_rel = np.zeros(n_s)                       # state difference just before the first divergence
    drift_frame = np.full(n_s, -1, np.int32)      # first update after which the states differ by more than 1e-9 (relative)
    drift_epi = np.full(n_s, np.nan)              # distance (px) of the pixel from the epipole of that update's frame
    xx = np.arange(b, w - b, dtype=np.float64)[None, :]
    yy = ys.astype(np.float64)[:, None]
    worst_flag = 0.0
    per_frame = []
    dg_prev = np.full(n_s, 3.0)
    cg_prev = np.full(n_s, 3.0)
    do_prev = np.full(n_s, 3.0)
    co_prev = np.full(n_s, 3.0)
    for i in range(1, F):
        T = seq.T_C_R(i)
        fl_o = np.zeros((h, w), np.uint8)
        k_o = np.zeros((h, w), np.int32)
        n_o = np.zeros((h, w), np.float64)
        oracle.update_ex(p, frames[0], frames[i], T.q, T.t, d_o, c_o, rows=(int(ys[0]), int(ys[-1]) + 1), row_stride=stride,
                         flags=fl_o, dbg_k=k_o, dbg_ncc64=n_o)
        if not a.fake_gpu:
            f.update(frames[i], T)
            fl_g = f.flags()
            ncc_g, trips_g, best_g = f.debug()
            d_g, c_g = f.download_state()
        else:
            fl_g = np.zeros((h, w), np.uint8)
            kk = np.zeros((h, w), np.int32)
            nn = np.zeros((h, w), np.float64)
            oracle.update_ex(p, frames[0], frames[i], T.q, T.t, d_g, c_g, rows=(int(ys[0]), int(ys[-1]) + 1), row_stride=stride,
                             flags=fl_g, dbg_k=kk, dbg_ncc64=nn)
            trips_g, best_g, ncc_g = kk >> 16, np.where((kk & 0xFFFF) == 0xFFFF, -1, kk & 0xFFFF), nn.astype(np.float32)
        FG, FO = fl_g[ys][:, xs], fl_o[ys][:, xs]
        TG, TO = trips_g[ys][:, xs], (k_o >> 16)[ys][:, xs]
        BO = (k_o & 0xFFFF)[ys][:, xs]
        BO = np.where(BO == 0xFFFF, -1, BO)
        BG = best_g[ys][:, xs]
        NG, NO = ncc_g[ys][:, xs], n_o[ys][:, xs]
        act = ((FG | FO) & 1) != 0
        gate = (FG & 1) != (FO & 1)
        both = ((FG & FO) & 1) != 0
        trip = both & (TG != TO)
        amax = both & ~trip & (BG != BO)
        acc = both & ~trip & ~amax & ((FG & 2) != (FO & 2))
        cls = np.where(gate, 1, np.where(trip, 2, np.where(amax, 3, np.where(acc, 4, 0)))).astype(np.int8)
        new = (cls != 0) & (first_frame < 0)
        if new.any():
            first_frame[new] = i
            first_class[new] = cls[new]
            rel = np.abs(dg_prev - do_prev) / np.maximum(np.abs(do_prev), 1e-300)
            relc = np.abs(cg_prev - co_prev) / np.maximum(np.abs(co_prev), 1e-300)
            pre_rel[new] = np.maximum(rel, relc)[new]
            for (r, c) in zip(*np.nonzero(new)):
                if len(first_info) < 4000:
                    first_info[(int(r), int(c))] = {
                        "y": int(ys[r]), "x": int(c + b), "update": i, "class": int(cls[r, c]),
                        "flags_gpu": int(FG[r, c]), "flags_ref": int(FO[r, c]), "trips_gpu": int(TG[r, c]), "trips_ref": int(TO[r, c]),
                        "best_gpu": int(BG[r, c]), "best_ref": int(BO[r, c]), "ncc_gpu_f32": float(NG[r, c]), "ncc_ref": float(NO[r, c]),
                        "pre_depth_gpu": float(dg_prev[r, c]), "pre_depth_ref": float(do_prev[r, c]),
                        "pre_cov2_gpu": float(cg_prev[r, c]), "pre_cov2_ref": float(co_prev[r, c])}
        mism = float((FG != FO).mean())
        worst_flag = max(worst_flag, mism)
        dg_prev, cg_prev = d_g[ys][:, xs].copy(), c_g[ys][:, xs].copy()
        do_prev, co_prev = d_o[ys][:, xs].copy(), c_o[ys][:, xs].copy()
        both_nan = np.isnan(dg_prev) & np.isnan(do_prev)
        ok3 = (np.abs(dg_prev - do_prev) <= 1e-3 * np.abs(do_prev)) | both_nan
        # numeric drift of the state while every decision so far was equal: where is the pixel relative to the epipole?
        def _rel(a, r):
            v = np.abs(a - r) / np.maximum(np.abs(r), 1e-300)
            return np.where(np.isnan(a) & np.isnan(r), 0.0, np.where(np.isnan(v), np.inf, v))
        rel_now = np.maximum(_rel(dg_prev, do_prev), _rel(cg_prev, co_prev))
        newd = (rel_now > 1e-9) & (drift_frame < 0) & ((first_frame < 0) | (first_frame == i))
        if newd.any():
            Ti = T.inverse()
            ex = p.fx * Ti.t[0] / Ti.t[2] + p.cx if Ti.t[2] != 0 else np.inf
            ey = p.fy * Ti.t[1] / Ti.t[2] + p.cy if Ti.t[2] != 0 else np.inf
            dist = np.sqrt((xx - ex) ** 2 + (yy - ey) ** 2)
            drift_frame[newd] = i
            drift_epi[newd] = dist[newd]
        per_frame.append({"update": i, "flag_mismatch": mism, "active_frac": float(act.mean()), "new_divergences": int(new.sum()),
                          "depth_within_1e-3": float(ok3.mean())})
        if i <= 3 or i % 25 == 0 or i == F - 1:
            print(f"[diag] update {i}: flag mismatch {mism:.2e}, active {act.mean():.3f}, new first-divergences {int(new.sum())}, "
                  f"depth within 1e-3 {ok3.mean():.5f} ({time.time() - t0:.0f}s)", flush=True)

    both_nan = np.isnan(dg_prev) & np.isnan(do_prev)
    rel_d = np.abs(dg_prev - do_prev) / np.maximum(np.abs(do_prev), 1e-300)
    rel_d[both_nan] = 0
    bad = ~((rel_d <= 1e-3) | both_nan)
    names = {0: "drift (no decision ever differed)", 1: "gate", 2: "trip", 3: "argmax", 4: "accept"}
    cls_f = lambda c: np.where(np.isnan(c), 3, np.where(c < p.min_cov, 0, np.where(c > p.max_cov, 1, 2)))
    out = {
        "workload": a.workload, "width": w, "height": h, "updates": F - 1, "rows": int(len(ys)), "row_stride": int(stride),
        "pixels": int(bad.size),
        "final_depth_within_1e-3": float(1.0 - bad.mean()),
        "final_depth_within_1e-6": float(((rel_d <= 1e-6) | both_nan).mean()),
        "final_class_mismatch": float((cls_f(cg_prev) != cls_f(co_prev)).mean()),
        "worst_per_update_flag_mismatch": worst_flag,
        "pixels_with_any_decision_difference": int((first_frame >= 0).sum()),
        "first_divergence_class_all": {names[k]: int(((first_class == k) & (first_frame >= 0)).sum()) for k in (1, 2, 3, 4)},
        "deviating_pixels(>1e-3)": int(bad.sum()),
        "deviating_by_first_divergence_class": {names[k]: int((bad & (first_class == k) & ((first_frame >= 0) | (k == 0))).sum())
                                                for k in (0, 1, 2, 3, 4)},
        "pre_divergence_state_rel_diff": {"median": float(np.median(pre_rel[first_frame >= 0])) if (first_frame >= 0).any() else None,
                                          "max": float(pre_rel[first_frame >= 0].max()) if (first_frame >= 0).any() else None},
        "first_divergence_update_histogram": {str(k): int(v) for k, v in zip(*np.unique(first_frame[first_frame >= 0] // 25 * 25, return_counts=True))},
    }
    drifted = drift_frame >= 0
    decided_first = (first_frame >= 0) & (~drifted | (first_frame <= drift_frame))
    out["root_cause"] = {
        "pixels_whose_state_drifted_>1e-9_before_or_without_any_decision_difference": int((drifted & ~decided_first).sum()),
        "pixels_whose_first_difference_was_a_decision_with_states_equal_to_1e-9": int(decided_first.sum()),
        "deviating(>1e-3)_by_root": {"state_drift_first": int((bad & drifted & ~decided_first).sum()), "decision_first": int((bad & decided_first).sum()),
                                     "neither": int((bad & ~drifted & (first_frame < 0)).sum())},
        "epipole_distance_px_at_first_drift": ({"median": float(np.nanmedian(drift_epi[drifted & ~decided_first])),
                                                "p90": float(np.nanquantile(drift_epi[drifted & ~decided_first], 0.9)),
                                                "max": float(np.nanmax(drift_epi[drifted & ~decided_first]))}
                                               if (drifted & ~decided_first).any() else None),
        "max_state_rel_diff_among_pixels_without_any_difference_flag": float(np.nanmax(np.where(~drifted & (first_frame < 0), rel_d, 0.0))),
    }
    # near-tie evidence for the argmax class: |ncc_gpu - ncc_ref| of the two different winners
    am = [v for v in first_info.values() if v["class"] == 3]
    if am:
        gaps = np.array([abs(v["ncc_gpu_f32"] - v["ncc_ref"]) for v in am])
        out["argmax_winner_ncc_gap"] = {"n": len(am), "median": float(np.median(gaps)), "p90": float(np.quantile(gaps, 0.9)), "max": float(gaps.max()),
                                        "adjacent_samples_frac": float(np.mean([abs(v["best_gpu"] - v["best_ref"]) == 1 for v in am]))}
    ac = [v for v in first_info.values() if v["class"] == 4]
    if ac:
        out["accept_ncc_minus_thresh"] = {"n": len(ac), "max_abs": float(max(abs(v["ncc_ref"] - p.ncc_thresh) for v in ac))}
    ex = sorted(first_info.values(), key=lambda v: v["update"])
    out["examples"] = ex[: a.examples]
    out["per_update"] = per_frame[:: max(1, len(per_frame) // 40)]
    txt = json.dumps(out, indent=1)
    print(txt)
    if a.out:
        Path(a.out).parent.mkdir(parents=True, exist_ok=True)
        Path(a.out).write_text(txt)


if __name__ == "__main__":
    main()
