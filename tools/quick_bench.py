"""Device-timed update sequence (development tool): python tools/quick_bench.py [workload] [frames] [reps]"""
import sys, time
sys.path.insert(0, ".")
import torch
from slamplay_b200.synth import make_sequence
from slamplay_b200.depth_filter import DepthFilter
wl = sys.argv[1] if len(sys.argv) > 1 else "hd_1920x1080"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 300
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
seq = make_sequence(wl, n_frames=n)
h, w = seq.shape
pitch = (w + 15) // 16 * 16
frames = torch.zeros((n, h, pitch), dtype=torch.uint8, device="cuda")
for i in range(n):
    seq.render_device(i, frames[i].data_ptr(), pitch, stream=torch.cuda.current_stream().cuda_stream)
torch.cuda.synchronize()
f = DepthFilter(seq.params, device=0)
f.set_reference_device(frames[0].data_ptr(), pitch)
poses = [seq.T_C_R(i) for i in range(n)]
st = torch.cuda.ExternalStream(f.stream())
best = 1e9
for r in range(reps + 1):
    f.fill_state(3.0, 3.0); f.counters(reset=True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for i in range(1, n):
        f.update_device(frames[i].data_ptr(), pitch, poses[i])
    e1.record(st); f.sync()
    ms = e0.elapsed_time(e1); c = f.counters()
    if r > 0: best = min(best, ms)
print(f"{wl} frames={n}: best {best:.2f} ms/seq  {c['interior']/best/1e6:.3f} G px-upd/s  {c['ncc_evals']/best/1e6:.2f} G NCC/s  evals={c['ncc_evals']} accepted={c['accepted']}")
