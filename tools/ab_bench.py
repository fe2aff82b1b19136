"""A/B of kernel variants in one process (development tool):
    python tools/ab_bench.py <workload> <frames> <reps> lib1.so lib2.so ...

One line per library: device-timed sequence, per-kernel times of one instrumented pass, and the SHA-256 of the final
maps (all variants must print the same digest: the arithmetic does not depend on scheduling choices).
"""
import hashlib, os, sys
sys.path.insert(0, ".")
import torch
from slamplay_b200 import _lib
from slamplay_b200.synth import make_sequence
from slamplay_b200.depth_filter import DepthFilter

wl, n, reps = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
libs = sys.argv[4:] or ["slamplay_b200/libdmf.so"]
seq = make_sequence(wl, n_frames=n)
h, w = seq.shape
pitch = (w + 15) // 16 * 16
frames = torch.zeros((n, h, pitch), dtype=torch.uint8, device="cuda")
for i in range(n):
    seq.render_device(i, frames[i].data_ptr(), pitch, stream=torch.cuda.current_stream().cuda_stream)
torch.cuda.synchronize()
poses = [seq.T_C_R(i) for i in range(n)]
for spec in libs:
    # "path.so" or "path.so:VAR=value[,VAR2=value2]" (environment read by the library at context creation, e.g. DMF_MOMENTS=legacy)
    lib, _, envs = spec.partition(":")
    for kv in filter(None, envs.split(",")):
        k, _, v = kv.partition("=")
        os.environ[k] = v
    os.environ["DMF_LIB"] = os.path.abspath(lib)
    _lib._cache.pop("dmf", None)
    try:
        # DMF_AB_CYCLIC=N: the context of rank 0 of an N-GPU run (1/N of the rows, whole-frame moment table)
        # DMF_AB_BLOCK_ROWS / DMF_AB_PART: rows per cyclic block (default 8) and which rank's share (default 0)
        ncyc = int(os.environ.get("DMF_AB_CYCLIC", "0"))
        brows, part = int(os.environ.get("DMF_AB_BLOCK_ROWS", "8")), int(os.environ.get("DMF_AB_PART", "0"))
        f = DepthFilter(seq.params, device=0, cyclic=(brows, ncyc, part)) if ncyc > 1 else DepthFilter(seq.params, device=0)
        f.set_reference_device(frames[0].data_ptr(), pitch)
        st = torch.cuda.ExternalStream(f.stream())
        best = 1e9
        for r in range(reps + 1):
            f.fill_state(3.0, 3.0); f.counters(reset=True)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(st)
            for i in range(1, n):
                f.update_device(frames[i].data_ptr(), pitch, poses[i])
            f.flush()
            e1.record(st); f.sync()
            ms = e0.elapsed_time(e1); c = f.counters()
            if r > 0: best = min(best, ms)
        d, v = f.download_state()
        dig = hashlib.sha256(d.tobytes() + v.tobytes()).hexdigest()[:16]
        f.fill_state(3.0, 3.0); f.set_timing(True)
        for i in range(1, n):
            f.update_device(frames[i].data_ptr(), pitch, poses[i])
        t = f.timing(reset=True)
        f.close()
        print(f"{os.path.basename(spec):40s} {wl} n={n}: {best:8.2f} ms  {c['interior']/best/1e6:6.3f} Gpx/s  {c['ncc_evals']/best/1e6:6.2f} GNCC/s | "
              f"setup {t['setup_ms']:.1f} mom {t['moments_ms']:.1f} ncc {t['ncc_ms']:.1f} fuse {t['fuse_ms']:.1f} ms | evals={c['ncc_evals']} sha={dig}", flush=True)
    except Exception as e:  # keep going: one broken variant must not waste the GPU call
        print(f"{os.path.basename(spec):24s} FAILED: {e!r}", flush=True)
    for kv in filter(None, envs.split(",")):
        os.environ.pop(kv.partition("=")[0], None)
