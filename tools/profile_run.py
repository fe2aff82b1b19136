"""Short single-GPU run of the update sequence for ncu (development tool).
    python tools/profile_run.py [workload] [n_frames] [update_to_bracket]
With update_to_bracket = K the work counters of update K alone are printed as one JSON line ("update_counters"), so that
the ncu counts of that update's ncc_kernel launch can be divided by its NCC evaluations."""
import json
import sys
sys.path.insert(0, ".")
import torch
from slamplay_b200.synth import make_sequence
from slamplay_b200.depth_filter import DepthFilter

wl = sys.argv[1] if len(sys.argv) > 1 else "hd_1920x1080"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
K = int(sys.argv[3]) if len(sys.argv) > 3 else -1
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
for i in range(1, n):
    if i == K:
        f.counters(reset=True)
    f.update_device(frames[i].data_ptr(), pitch, seq.T_C_R(i))
    if i == K:
        c = f.counters(reset=True)
        print("update_counters " + json.dumps({"workload": wl, "update": K, **c}), flush=True)
f.sync()
print(f.counters())
