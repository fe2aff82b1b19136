"""Per-update device time of a sequence, whole frame vs one context of an N-rank block-cyclic split (development tool):
    python tools/per_update_times.py <workload> <frames> <n_ranks> [part]
Prints, for groups of updates, the time of the whole-frame context / N (the ideal share) beside the time of the split
context: where along the sequence strong scaling is lost (fixed per-update costs vs work)."""
import sys
sys.path.insert(0, ".")
import numpy as np
import torch
from slamplay_b200.synth import make_sequence
from slamplay_b200.depth_filter import DepthFilter

wl, n, ranks = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
part = int(sys.argv[4]) if len(sys.argv) > 4 else 0
seq = make_sequence(wl, n_frames=n)
h, w = seq.shape
pitch = (w + 15) // 16 * 16
frames = torch.zeros((n, h, pitch), dtype=torch.uint8, device="cuda")
for i in range(n):
    seq.render_device(i, frames[i].data_ptr(), pitch, stream=torch.cuda.current_stream().cuda_stream)
torch.cuda.synchronize()
poses = [seq.T_C_R(i) for i in range(n)]


def run(f):
    f.set_reference_device(frames[0].data_ptr(), pitch)
    st = torch.cuda.ExternalStream(f.stream())
    best = None
    for rep in range(3):
        f.fill_state(3.0, 3.0)
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(n)]
        ev[0].record(st)
        for i in range(1, n):
            f.update_device(frames[i].data_ptr(), pitch, poses[i])
            ev[i].record(st)
        f.flush(); f.sync()
        t = np.array([ev[i - 1].elapsed_time(ev[i]) for i in range(1, n)])
        best = t if best is None else np.minimum(best, t)
    f.close()
    return best


whole = run(DepthFilter(seq.params, device=0))
split = run(DepthFilter(seq.params, device=0, cyclic=(8, ranks, part)))
print(f"{wl} {n} frames, {ranks} ranks, part {part}: whole {whole.sum():.2f} ms (/{ranks} = {whole.sum() / ranks:.2f}), split {split.sum():.2f} ms")
edges = [0, 1, 2, 4, 8, 16, 32, 64, 100, 150, 200, 250, 300, 400, n - 1]
for a, b in zip(edges, edges[1:]):
    if a >= n - 1:
        break
    b = min(b, n - 1)
    wi, sp = whole[a:b].sum() / ranks, split[a:b].sum()
    print(f"updates {a + 1:3d}..{b:3d}: ideal {1e3 * wi / (b - a):8.1f} us/update   split {1e3 * sp / (b - a):8.1f} us/update   lost {sp - wi:6.2f} ms")
