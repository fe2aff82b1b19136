"""torchrun --nproc-per-node N tools/multi_gpu_check.py [workload] [frames]
Row-band sharded run over N GPUs vs a single-GPU run on rank 0: the gathered maps must be bit-identical."""
import os, sys, time
sys.path.insert(0, ".")
import numpy as np, torch, torch.distributed as dist
from slamplay_b200.synth import make_sequence
from slamplay_b200.depth_filter import DepthFilter
from slamplay_b200.sharded import ShardedDepthFilter

wl = sys.argv[1] if len(sys.argv) > 1 else "hd_1920x1080"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 30
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
seq = make_sequence(wl, n_frames=n)
h, w = seq.shape
pitch = (w + 15) // 16 * 16
frames = None
if rank == 0:
    frames = torch.zeros((n, h, pitch), dtype=torch.uint8, device=dev)
    for i in range(n):
        seq.render_device(i, frames[i].data_ptr(), pitch, stream=torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
sf = ShardedDepthFilter(seq.params, device=lr)
sf.set_reference(frames[0] if rank == 0 else None)
poses_all = [seq.T_C_R(i) for i in range(n)]
for rep in range(2):
    sf.fill_state(3.0, 3.0)
    sf.counters(reset=True)
    dist.barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    poses = sf.broadcast_poses(poses_all if rank == 0 else None)
    for i in range(1, n):
        sf.update(frames[i] if rank == 0 else None, poses[i])
    res = sf.gather_state()
    dist.barrier(); torch.cuda.synchronize()
    dt = time.perf_counter() - t0
cnt = sf.counters()
if rank == 0:
    d, c = res[0].cpu().numpy(), res[1].cpu().numpy()
    f = DepthFilter(seq.params, device=lr)
    f.set_reference_device(frames[0].data_ptr(), pitch)
    f.fill_state(3.0, 3.0)
    for i in range(1, n):
        f.update_device(frames[i].data_ptr(), pitch, poses_all[i])
    d1, c1 = f.download_state()
    c1cnt = f.counters()
    same = np.array_equal(d, d1, equal_nan=True) and np.array_equal(c, c1, equal_nan=True)
    print(f"world={world} {wl} frames={n}: sharded {dt*1e3:.1f} ms, bit-identical to 1 GPU: {same}, counters equal: "
          f"{all(cnt[k] == c1cnt[k] for k in ('interior','active','ncc_evals','accepted'))}", flush=True)
    assert same
dist.destroy_process_group()
