"""Histogram of epipolar trip counts per active pixel over a sequence (development tool)."""
import sys
sys.path.insert(0, ".")
import numpy as np, torch
from slamplay_b200.synth import make_sequence
from slamplay_b200.depth_filter import DepthFilter
wl = sys.argv[1] if len(sys.argv) > 1 else "hd_1920x1080"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 300
seq = make_sequence(wl, n_frames=n)
h, w = seq.shape
pitch = (w + 15) // 16 * 16
frames = torch.zeros((n, h, pitch), dtype=torch.uint8, device="cuda")
for i in range(n):
    seq.render_device(i, frames[i].data_ptr(), pitch, stream=torch.cuda.current_stream().cuda_stream)
torch.cuda.synchronize()
f = DepthFilter(seq.params, device=0)
f.set_reference_device(frames[0].data_ptr(), pitch)
f.fill_state(3.0, 3.0)
f.enable_flags(True)
edges = [0, 1, 5, 9, 17, 33, 65, 129, 257, 400]
for i in range(1, n):
    f.update_device(frames[i].data_ptr(), pitch, seq.T_C_R(i))
    if i in (1, 2, 3, 5, 10, 20, 40, 80, 150, 250, 299, 499):
        ncc, trips, k = f.debug()
        fl = f.flags()
        t = trips[20:-20, 20:-20].ravel()
        act = (fl[20:-20, 20:-20].ravel() & 1) > 0
        hist_px = np.histogram(t[act], bins=edges)[0]
        hist_w = np.histogram(t[act], bins=edges, weights=t[act])[0]
        tot = t[act].sum()
        print(f"frame {i}: active {act.mean():.3f} mean trips/active {t[act].mean():.1f} | px share by trips {edges[1:]}: "
              + " ".join(f"{x/act.sum():.2f}" for x in hist_px) + " | sample share: " + " ".join(f"{x/max(tot,1):.2f}" for x in hist_w))
