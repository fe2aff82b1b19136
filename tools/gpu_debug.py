import sys, numpy as np
sys.path.insert(0, ".")
import oracle
from slamplay_b200.synth import make_sequence
from slamplay_b200.depth_filter import DepthFilter
n = int(sys.argv[1]) if len(sys.argv) > 1 else 10
seq = make_sequence("remode_640x480", n_frames=n)
p = seq.params; h, w = seq.shape
frames = [seq.render_host(i) for i in range(n)]
f = DepthFilter(p); f.set_reference(frames[0]); f.fill_state(3.0, 3.0); f.enable_flags(True)
d_ref = np.full((h, w), 3.0); c_ref = np.full((h, w), 3.0)
for i in range(1, n):
    T = seq.T_C_R(i)
    # start every frame from the oracle's state so that differences do not compound
    f.upload_state(d_ref, c_ref)
    d_in, c_in = d_ref.copy(), c_ref.copy()
    f.update(frames[i], T)
    fl = f.flags(); gncc, gn, gk = f.debug()
    d_gpu, c_gpu = f.download_state()
    fl_ref = np.zeros((h, w), np.uint8); oncc = np.zeros((h, w), np.float32); on = np.zeros((h, w), np.int32)
    oracle.update(p, frames[0], frames[i], T.q, T.t, d_ref, c_ref, flags=fl_ref, dbg_ncc=oncc, dbg_n=on)
    I = (slice(20, h - 20), slice(20, w - 20))
    mm = fl[I] != fl_ref[I]
    dn = np.abs(gncc[I] - oncc[I])
    print(f"frame {i}: flag mismatch {mm.mean():.3e}  max|dNCC| {dn.max():.3e}  mean trip gpu {gn[I].mean():.2f}  evals oracle {on[I].mean():.2f}"
          f"  depth mismatch {(np.abs(d_gpu[I]-d_ref[I]) > 1e-6*np.abs(d_ref[I])).mean():.3e}")
    ys, xs = np.nonzero(dn > 1e-4)
    for y, x in list(zip(ys, xs))[:8]:
        Y, X = y + 20, x + 20
        print(f"   px({X},{Y}) mu={d_in[Y,X]:.6f} cov2={c_in[Y,X]:.3e} gpu ncc={gncc[Y,X]:.6f} trips={gn[Y,X]} k={gk[Y,X]} | oracle ncc={oncc[Y,X]:.6f} evals={on[Y,X]} flags gpu={fl[Y,X]} ref={fl_ref[Y,X]}")
