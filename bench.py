#!/usr/bin/env python
"""bench.py — pixel-updates/s of the dense monocular depth filter on 1/2/4/8 B200.

Metric (BASELINE.json): pixel-updates/s = interior pixels visited by the loops of update()
(dense_mapping/test_monocular_mapping.cpp:357,363) x frames after the reference / time.

A STEP is one pass of the hot path over one synthetic sequence: state reset to 3.0 / 3.0
(ref:270-278), then update() for every frame after the reference frame.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME] [--frames F]
  python bench.py --impl reference ...        # the reference's CPU path on the host cores
  torchrun --nproc-per-node N bench.py --gpus N ...   (N > 1: row-band sharding, NCCL)

Prints ONE JSON line (rank 0).  See DESIGN.md §5 for how every field is derived.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC = "pixel-updates/s (px x frames/s) of dense mono depth filter"
UNIT = "px-updates/s"
DEFAULT_WORKLOAD = "uhd_3840x2160"
FP32_PEAK_TFLOPS_NOMINAL = 148 * 128 * 2 * 1.965e9 / 1e12  # 74.4: 148 SMs x 128 FMA lanes x 2 x 1965 MHz
FLOP_PER_NCC, FLOP_PER_ACTIVE, FLOP_PER_ACCEPT = 600.0, 150.0, 300.0  # SURVEY.md §8d (definitional)


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def measured_peaks() -> dict:
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return json.loads(p.read_text())
        except Exception:
            pass
    return {}


def ncc_traffic(workload: str):
    """dram__bytes_read.sum + dram__bytes_write.sum of one ncc_kernel launch (bytes) from the committed ncu --set full
    capture (profiles/r01_ncc_traffic.json): a mid-sequence launch of the 1080p workload; null for other workloads."""
    p = ROOT / "profiles" / "r01_ncc_traffic.json"
    try:
        j = json.loads(p.read_text())
        if j.get("workload") == workload:
            return j["dram_bytes_read"] + j["dram_bytes_write"]
    except Exception:
        pass
    return None


class ClockSampler:
    """nvidia-smi clock / throttle-reason sampler running during the timed region."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap,timestamp")

    def __init__(self, gpu_index: int = 0):
        self.gpu = gpu_index
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50",
                                          "-i", str(self.gpu)], stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def mark(self, name: str) -> None:
        """Remember the wall-clock time of the start / end of the timed region."""
        setattr(self, name, time.time())

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        import datetime
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        t0, t1 = getattr(self, "t_begin", None), getattr(self, "t_end", None)
        for line in Path(self.path).read_text().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 10:
                continue
            try:
                if t0 is not None and t1 is not None:  # keep the samples taken DURING the timed region
                    ts = datetime.datetime.strptime(f[9], "%Y/%m/%d %H:%M:%S.%f").timestamp()
                    if ts < t0 - 0.05 or ts > t1 + 0.05:
                        continue
                sm.append(float(f[1])); mx.append(float(f[2])); pw.append(float(f[3]))
            except ValueError:
                continue
            for name, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        try:
            os.unlink(self.path)
        except OSError:
            pass
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "power_w_max": max(pw), "samples": len(sm),
                "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------
# sequences
def build_sequence(workload: str, frames: int | None):
    from slamplay_b200.synth import make_sequence

    return make_sequence(workload, n_frames=frames)


def render_frames_gpu(seq, torch, device):
    """All frames of the sequence rendered on the GPU into one (F, H, pitch) uint8 tensor."""
    h, w = seq.shape
    pitch = (w + 15) // 16 * 16
    frames = torch.zeros((seq.n_frames, h, pitch), dtype=torch.uint8, device=device)
    s = torch.cuda.current_stream().cuda_stream
    for i in range(seq.n_frames):
        seq.render_device(i, frames[i].data_ptr(), pitch, stream=s)
    torch.cuda.synchronize()
    return frames, pitch


# ------------------------------------------------------------------------------------------------
# the reference's CPU path (oracle port / compiled reference TU) on the host cores
def default_cpu_rows(seq, target_ncc: float) -> int:
    """Rows of the CPU sample so that one pass costs about `target_ncc` NCC evaluations (the CPU path does
    ~1.9 M NCC/s per core with the reference's heap allocations on the GPU box's host): ~20 NCC per pixel-update on these sequences."""
    p = seq.params
    per_row = (p.width - 2 * p.border) * 20.0 * (seq.n_frames - 1)
    return max(1, int(round(target_ncc / per_row)))


def cpu_rows_sample(p, n_rows: int):
    lo, hi = p.border, p.height - p.border
    n_rows = max(1, min(n_rows, hi - lo))
    stride = max(1, (hi - lo) // n_rows)
    rows = list(range(lo + stride // 2, hi, stride))[:n_rows]
    return rows, stride


def run_cpu_sequence(seq, host_frames, rows_spec, heap: bool, use_ref_tu: bool, threads: int):
    """Runs update() over the whole sequence on a row subset (pixels are independent, so a row subset
    is an exact sub-problem).  Returns (seconds, counters dict, depth, cov2)."""
    import oracle

    p = seq.params
    h, w = seq.shape
    first, stride, n_rows = rows_spec
    depth = np.full((h, w), 3.0)
    cov2 = np.full((h, w), 3.0)
    cnt = oracle.Counters()
    oracle.lib().dmo_set_threads(threads)
    poses = [seq.T_C_R(i) for i in range(seq.n_frames)]
    t0 = time.perf_counter()
    if use_ref_tu:
        for i in range(1, seq.n_frames):
            oracle.ref_update(host_frames[0], host_frames[i], poses[i].q, poses[i].t, depth, cov2)
        cnt.frames = seq.n_frames - 1
        cnt.interior = (seq.n_frames - 1) * (h - 2 * p.border) * (w - 2 * p.border)
    else:
        for i in range(1, seq.n_frames):
            oracle.update(p, host_frames[0], host_frames[i], poses[i].q, poses[i].t, depth, cov2,
                          rows=(first, first + stride * n_rows), row_stride=stride, heap=heap, counters=cnt)
    dt = time.perf_counter() - t0
    return dt, cnt.as_dict(), depth, cov2


def reference_arm(args) -> dict:
    """--impl reference: the reference's CPU implementation of update() on this box's host cores."""
    import oracle

    oracle.build(ref=True)
    seq = build_sequence(args.workload, args.frames)
    p = seq.params
    h, w = seq.shape
    cores = os.cpu_count() or 1
    use_ref_tu = (w, h) == (640, 480) and oracle.ref_lib() is not None and not args.force_port
    # frames: rendered on the CPU for small sizes, on the GPU when available for large ones
    t0 = time.perf_counter()
    host_frames = None
    try:
        import torch
        if torch.cuda.is_available() and w * h > 640 * 480:
            fr, pitch = render_frames_gpu(seq, torch, torch.device("cuda", 0))
            host_frames = [np.ascontiguousarray(fr[i, :, :w].cpu().numpy()) for i in range(seq.n_frames)]
            del fr
            torch.cuda.empty_cache()
    except Exception as e:  # pragma: no cover
        log("GPU rendering of the input frames unavailable:", e)
    if host_frames is None:
        host_frames = [seq.render_host(i) for i in range(seq.n_frames)]
    log(f"[reference] inputs ready in {time.perf_counter() - t0:.1f}s")
    if use_ref_tu:
        rows, stride = list(range(p.border, h - p.border)), 1
        sample = f"full {w}x{h} sequence, {seq.n_frames - 1} updates, compiled reference TU (oracle/_ref)"
    else:
        rows, stride = cpu_rows_sample(p, args.cpu_rows or default_cpu_rows(seq, 1.2e7 * cores))  # ~7 s per step
        sample = (f"{len(rows)} of {h - 2 * p.border} interior rows (every {stride}th from y={rows[0]}) x all "
                  f"{seq.n_frames - 1} updates of {args.workload}; oracle port with the reference's per-NCC heap allocations")
    spec = (rows[0], stride, len(rows))
    times, cnts = [], None
    for it in range(args.warmup + args.steps):
        dt, cnts, _, _ = run_cpu_sequence(seq, host_frames, spec, heap=True, use_ref_tu=use_ref_tu, threads=cores)
        log(f"[reference] step {it}: {dt:.2f}s")
        if it >= args.warmup:
            times.append(dt)
    t = sum(times) / len(times)
    value = cnts["interior"] / t
    out = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": t * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": args.workload, "width": w, "height": h, "frames": seq.n_frames, "sample": sample},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "reference" if use_ref_tu else "port",
                         "sample": sample, "ncc_evals_per_s": cnts["ncc_evals"] / t if cnts["ncc_evals"] else None},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    return out


# ------------------------------------------------------------------------------------------------
def ours(args) -> dict | None:
    import torch
    import torch.distributed as dist

    from slamplay_b200 import build as dmf_build
    from slamplay_b200.depth_filter import DepthFilter
    from slamplay_b200.sharded import ShardedDepthFilter, band_rows

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the depth-filter path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=device)
    if args.gpus != world and rank == 0:
        log(f"warning: --gpus {args.gpus} but WORLD_SIZE={world}; using {world}")
    if rank == 0:
        dmf_build.build_all()
    if world > 1:
        dist.barrier()

    seq = build_sequence(args.workload, args.frames)
    p = seq.params
    h, w = seq.shape
    F = seq.n_frames
    n_upd = F - 1
    interior_per_frame = (h - 2 * p.border) * (w - 2 * p.border)

    # inputs resident in HBM before the timed region (rank 0 holds the sequence)
    t0 = time.perf_counter()
    if rank == 0:
        frames, pitch = render_frames_gpu(seq, torch, device)
        log(f"rendered {F} frames {w}x{h} on the GPU in {time.perf_counter() - t0:.1f}s "
            f"({frames.numel() / 1e9:.2f} GB in HBM; inputs larger than L2)")
    else:
        frames, pitch = None, (w + 15) // 16 * 16
    poses_all = [seq.T_C_R(i) for i in range(F)]

    sf = ShardedDepthFilter(p, device=local_rank, layout=args.layout, block_rows=args.block_rows, n_ring=args.ring)
    sf.set_reference(frames[0] if rank == 0 else None)
    ctx_stream = sf.ctx_stream

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def one_step():
        """One full sequence; everything asynchronous."""
        sf.fill_state(3.0, 3.0)
        poses = sf.broadcast_poses(poses_all if rank == 0 else None) if world > 1 else [(T.q, T.t) for T in poses_all]
        if world > 1:
            # the frame of update i+1 is announced before update i is launched: its broadcast runs one update early
            sf.prefetch(frames[1] if rank == 0 else None)
            for i in range(1, F):
                if i + 1 < F:
                    sf.prefetch(frames[i + 1] if rank == 0 else None)
                sf.update(None, poses[i])
            return sf.gather_state()
        for i in range(1, F):
            sf.update(frames[i], poses[i])
        if world > 1:
            return sf.gather_state()
        sf.flush()  # the deferred fusion of the last update belongs to this step
        return None

    # ---- device-timed value ---------------------------------------------------------------
    for _ in range(args.warmup):
        one_step()
    sf.counters(reset=True)
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    sampler.mark("t_begin")
    ev0.record(ctx_stream)
    for _ in range(args.steps):
        one_step()
    ev1.record(ctx_stream)
    barrier()
    sampler.mark("t_end")
    ms_total = ev0.elapsed_time(ev1)
    clocks = sampler.stop() if rank == 0 else None
    if world > 1:
        t = torch.tensor([ms_total], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
    cnt = sf.counters(reset=True)
    # one extra, untimed-for-the-metric step with CUDA events around every kernel: the share and the average
    # launch duration of the dominant kernel (ncc_kernel) for the roofline block
    sf.filter.set_timing(True)
    t_host0 = time.perf_counter()
    one_step()
    t_enqueue = time.perf_counter() - t_host0  # host time to enqueue + finish one instrumented step
    ktime = sf.filter.timing(reset=True)
    sf.filter.set_timing(False)
    sf.counters(reset=True)
    per_rank = None
    if world > 1:
        mine = torch.tensor([ktime["setup_ms"], ktime["moments_ms"], ktime["ncc_ms"], ktime["fuse_ms"], t_enqueue * 1e3],
                            dtype=torch.float64, device=device)
        allr = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(allr, mine)
        per_rank = [[round(float(v), 2) for v in t.tolist()] for t in allr]
    ms_step = ms_total / args.steps
    px_updates = interior_per_frame * n_upd
    value = px_updates / (ms_step * 1e-3)

    result = None
    if rank == 0:
        ncc = cnt["ncc_evals"] / args.steps
        act = cnt["active"] / args.steps
        acc = cnt["accepted"] / args.steps
        flops_step = FLOP_PER_NCC * ncc + FLOP_PER_ACTIVE * act + FLOP_PER_ACCEPT * acc
        k_frames = max(ktime["frames"], 1)
        ncc_launch_ms = ktime["ncc_ms"] / k_frames
        k_total = ktime["moments_ms"] + ktime["setup_ms"] + ktime["ncc_ms"] + ktime["fuse_ms"]
        ncc_flops_launch = FLOP_PER_NCC * ncc / n_upd / world  # algorithmic FP32 flops of one ncc_kernel launch on one GPU
        peaks = measured_peaks()
        hbm_peak = peaks.get("hbm_gbs", 6650.0)
        # compulsory HBM bytes per step: frame + cov read + depth read of active + writes of accepted
        hbm_bytes = n_upd * (w * h + 8 * interior_per_frame / world) + 8 * act + 16 * acc
        result = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u8 x u8 -> s32 (dp4a) NCC moments, f64 NCC combine / geometry / fusion",
            "data": "synthetic",
            "config": {"workload": args.workload, "width": w, "height": h, "frames": F, "updates_per_step": n_upd,
                       "interior_px_per_frame": interior_per_frame, "init_depth": 3.0, "init_cov2": 3.0,
                       "ncc_window": "7x7", "parallelism": (f"{args.layout} row blocks x{world}" + (f" ({args.block_rows} rows)" if args.layout == "cyclic" else "")) if world > 1 else "single GPU",
                       "l2_policy": f"inputs larger than L2 ({F * h * pitch / 1e6:.0f} MB of frames per step)",
                       "ncc_evals_per_step": ncc, "active_px_per_step": act, "accepted_per_step": acc},
            "ncc_evals_per_s": ncc / (ms_step * 1e-3),
            "roofline": {
                "bound": "fp32-issue (SURVEY.md 8d: the path is neither HBM- nor tensor-bound)",
                "kernel": "dmf::ncc_kernel", "avg_launch_ms": ncc_launch_ms,
                "achieved": ncc_flops_launch / (ncc_launch_ms * 1e-3) / 1e12,
                "peak": FP32_PEAK_TFLOPS_NOMINAL, "unit": "TFLOP/s",
                "frac": ncc_flops_launch / (ncc_launch_ms * 1e-3) / 1e12 / FP32_PEAK_TFLOPS_NOMINAL,
                "traffic": ncc_traffic(args.workload),
                "traffic_note": "DRAM bytes of the ncc_kernel launch of frame 40 (profiles/r01_ncc_traffic.json); null where no capture exists",
                "peak_source": "nominal FP32 FMA peak (148 SM x 128 lanes x 2 x 1965 MHz); MEASURED_PEAKS.json has no FP32 figure",
                "flop_model": "600 algorithmic FP32 flop per NCC evaluation (SURVEY.md 8d); the kernel itself computes the NCC from exact "
                              "integer moments (IDP.4A + a per-frame moment table) and is bound by L1/TEX gathers, see profiles/",
                "kernel_share_of_step": ktime["ncc_ms"] / k_total if k_total else None,
                "kernel_ms_per_step": {k: ktime[k] for k in ("moments_ms", "setup_ms", "ncc_ms", "fuse_ms")},
                "per_rank_ms[setup,moments,ncc,fuse,host_step]": per_rank,
                "whole_step": {"achieved": flops_step / (ms_step * 1e-3) / 1e12 / world,
                               "frac": flops_step / (ms_step * 1e-3) / 1e12 / world / FP32_PEAK_TFLOPS_NOMINAL,
                               "flop_model": "600/NCC + 150/active px + 300/accepted px, per GPU"},
                "hbm": {"achieved": hbm_bytes / (ms_step * 1e-3) / 1e9 / world, "peak": hbm_peak, "unit": "GB/s",
                        "frac": hbm_bytes / (ms_step * 1e-3) / 1e9 / world / hbm_peak,
                        "peak_source": "measured (MEASURED_PEAKS.json)" if "hbm_gbs" in peaks else "fallback"},
            },
            "clocks": clocks,
            # per step: state fill, setup_kernel of the first update, advance_kernel of the others (fusion of the
            # previous update + setup), moments_kernel and ncc_kernel per update, fuse_kernel of the last update
            "gpu_launches": args.steps * (3 * n_upd + 2),
        }

    # ---- end-to-end through the public API with HOST buffers ---------------------------------
    if not args.no_e2e:
        if world == 1:
            result["e2e"] = e2e_run(args, seq, frames, pitch, torch, device)
        else:
            e2e = e2e_run_sharded(args, seq, frames, sf, torch, dist, rank, world, device, poses_all)
            if rank == 0:
                result["e2e"] = e2e
    if world > 1:
        dist.barrier()

    # ---- N > 1: the gathered maps must be bit-identical to a single-GPU run (ref:366,546-564: pixels are independent)
    if world > 1 and not args.no_parity:
        gathered = one_step()
        if rank == 0:
            result["parity_sample"] = multi_gpu_parity(seq, frames, pitch, gathered, poses_all, torch, local_rank)
        dist.barrier()

    # ---- CPU baseline (rank 0, N == 1 only) + parity on the sampled rows -------------------
    if rank == 0 and world == 1 and not args.no_cpu:
        try:
            result.update(cpu_baseline_and_parity(args, seq, frames, sf, torch))
        except Exception as e:  # the baseline must never sink the bench line
            result["cpu_baseline"] = {"error": repr(e)}
    sf.close()
    if world > 1:
        dist.destroy_process_group()
    return result


def e2e_run(args, seq, frames, pitch, torch, device) -> dict:
    """Same metric through the host-facing API: every update() H2D-copies its frame from pinned host
    memory (double-buffered against the previous kernel), the maps are read back at the end of the step."""
    from slamplay_b200.depth_filter import DepthFilter

    p = seq.params
    h, w = seq.shape
    F = seq.n_frames
    host = torch.empty((F, h, w), dtype=torch.uint8, pin_memory=True)
    host.copy_(frames[:, :, :w])
    torch.cuda.synchronize()
    depth = torch.empty((h, w), dtype=torch.float64, pin_memory=True).numpy()
    cov2 = torch.empty((h, w), dtype=torch.float64, pin_memory=True).numpy()
    poses = [seq.T_C_R(i) for i in range(F)]
    f = DepthFilter(p, device=device.index)
    ref_np = host[0].numpy()
    f.set_reference(ref_np)
    base = host.data_ptr()
    fb = h * w

    def step():
        f.fill_state(3.0, 3.0)
        for i in range(1, F):
            f.update_ptr(base + i * fb, w, poses[i])
        f.download_state(depth, cov2)

    for _ in range(min(args.warmup, 2)):
        step()
    f.sync()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    f.sync()
    dt = (time.perf_counter() - t0) / args.steps
    f.close()
    interior = (h - 2 * p.border) * (w - 2 * p.border) * (F - 1)
    return {"value": interior / dt, "unit": UNIT, "h2d_bytes_per_step": (F - 1) * h * w, "d2h_bytes_per_step": 16 * h * w,
            "ms_per_step": dt * 1e3, "api": "DepthFilter.update (dmf_update, pinned host frames) + download_state"}


def e2e_run_sharded(args, seq, frames, sf, torch, dist, rank, world, device, poses_all) -> dict:
    """N > 1: rank 0 holds the frames in pinned host memory; per update H2D on rank 0 -> NCCL broadcast ->
    band update on every rank; per step gather of the bands and D2H of both maps on rank 0."""
    p = seq.params
    h, w = seq.shape
    F = seq.n_frames
    host = None
    if rank == 0:
        host = torch.empty((F, h, w), dtype=torch.uint8, pin_memory=True)
        host.copy_(frames[:, :, :w])
        out_d = torch.empty((h, w), dtype=torch.float64, pin_memory=True)
        out_c = torch.empty((h, w), dtype=torch.float64, pin_memory=True)
    torch.cuda.synchronize()

    def step():
        sf.fill_state(3.0, 3.0)
        poses = sf.broadcast_poses(poses_all if rank == 0 else None)
        sf.prefetch_host(host[1] if rank == 0 else None)
        for i in range(1, F):
            if i + 1 < F:
                sf.prefetch_host(host[i + 1] if rank == 0 else None)
            sf.update_host(None, poses[i])
        res = sf.gather_state()
        if rank == 0:
            out_d.copy_(res[0], non_blocking=True)
            out_c.copy_(res[1], non_blocking=True)
            torch.cuda.current_stream().synchronize()

    for _ in range(min(args.warmup, 2)):
        step()
    dist.barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dist.barrier(); torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / args.steps
    t = torch.tensor([dt], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dt = float(t.item())
    interior = (h - 2 * p.border) * (w - 2 * p.border) * (F - 1)
    return {"value": interior / dt, "unit": UNIT, "h2d_bytes_per_step": (F - 1) * h * w, "d2h_bytes_per_step": 16 * h * w,
            "ms_per_step": dt * 1e3, "api": "ShardedDepthFilter.update_host (pinned host frames on rank 0, NCCL broadcast) + gather_state + D2H"}


def multi_gpu_parity(seq, frames, pitch, gathered, poses_all, torch, device_index) -> dict:
    """Rank 0: the same sequence on ONE context; SHA-256 and bitwise comparison with the maps gathered from N ranks."""
    import hashlib

    from slamplay_b200.depth_filter import DepthFilter

    h, w = seq.shape
    multi = [t.cpu().numpy().copy() for t in gathered]
    f = DepthFilter(seq.params, device=device_index)
    f.set_reference_device(frames[0].data_ptr(), pitch)
    f.fill_state(3.0, 3.0)
    s = torch.cuda.current_stream().cuda_stream
    for i in range(1, seq.n_frames):
        f.update_device(frames[i].data_ptr(), pitch, poses_all[i], wait_stream=s)
    single = f.download_state()
    f.close()
    b = seq.params.border
    I = (slice(b, h - b), slice(b, w - b))
    sha = lambda a: hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()
    same = [float((m[I].view(np.int64) == s1[I].view(np.int64)).mean()) for m, s1 in zip(multi, single)]
    return {"against": "the same sequence on one GPU (rank 0), all interior pixels",
            "sha256_depth": sha(multi[0][I]), "sha256_depth_1gpu": sha(single[0][I]),
            "sha256_cov2": sha(multi[1][I]), "sha256_cov2_1gpu": sha(single[1][I]),
            "bit_identical": bool(same[0] == 1.0 and same[1] == 1.0), "bitwise_equal_frac": {"depth": same[0], "cov2": same[1]}}


def cpu_baseline_and_parity(args, seq, frames, sf, torch) -> dict:
    import oracle

    oracle.build(ref=False)
    p = seq.params
    h, w = seq.shape
    cores = os.cpu_count() or 1
    rows, stride = cpu_rows_sample(p, args.cpu_rows or default_cpu_rows(seq, 3.6e7 * cores))  # ~15-20 s (measured: ~1.9 M NCC/s per host core)
    host_frames = frames[:, :, :w].cpu().numpy()
    spec = (rows[0], stride, len(rows))
    dt, cnts, d_ref, c_ref = run_cpu_sequence(seq, host_frames, spec, heap=True, use_ref_tu=False, threads=cores)
    sample = (f"{len(rows)} of {h - 2 * p.border} interior rows (every {stride}th from y={rows[0]}) x all {seq.n_frames - 1} "
              f"updates; oracle port with the reference's per-NCC heap allocations, {cores} OpenMP threads")
    out = {"cpu_baseline": {"value": cnts["interior"] / dt, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
                            "seconds": dt, "ncc_evals_per_s": cnts["ncc_evals"] / dt}}
    # parity of the GPU maps (state left by the last timed step) on exactly those rows
    sf.filter.sync()
    d_gpu = sf.depth_t.cpu().numpy()
    c_gpu = sf.cov2_t.cpu().numpy()
    ys = np.array(rows)
    xs = slice(p.border, w - p.border)
    dg, dr, cg, cr = d_gpu[ys, xs], d_ref[ys, xs], c_gpu[ys, xs], c_ref[ys, xs]
    both_nan = np.isnan(dg) & np.isnan(dr)
    ok = (np.abs(dg - dr) <= 1e-3 * np.abs(dr)) | both_nan
    cls = lambda c: np.where(np.isnan(c), 3, np.where(c < p.min_cov, 0, np.where(c > p.max_cov, 1, 2)))
    out["parity_sample"] = {"rows": len(rows), "depth_within_1e-3": float(ok.mean()),
                            "final_class_mismatch": float((cls(cg) != cls(cr)).mean()),
                            "converged_frac_ref": float((cr < p.min_cov).mean())}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD)
    ap.add_argument("--frames", type=int, default=None, help="override the number of frames (incl. the reference frame)")
    ap.add_argument("--cpu-rows", type=int, default=None, help="rows of the CPU-baseline sample (default: host cores)")
    ap.add_argument("--layout", default="cyclic", choices=["cyclic", "bands"], help="row ownership for N > 1")
    ap.add_argument("--block-rows", type=int, default=8, help="rows per block of the cyclic layout")
    ap.add_argument("--ring", type=int, default=3, help="frames in flight per rank for N > 1 (broadcast ring depth)")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-parity", action="store_true", help="N > 1: skip the bitwise comparison with a single-GPU run")
    ap.add_argument("--force-port", action="store_true", help="--impl reference: use the oracle port even at 640x480")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)

    # The contract is ONE JSON line on stdout.  Native libraries print there too (NCCL: "NCCL version ..."), so file
    # descriptor 1 is pointed at stderr for the whole run and the JSON line goes to a saved duplicate of the real stdout.
    sys.stdout.flush()
    real_stdout = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)

    if args.impl == "reference":
        if int(os.environ.get("RANK", "0")) != 0:
            return 0
        print(json.dumps(reference_arm(args)), file=real_stdout, flush=True)
        return 0
    res = ours(args)
    if res is not None:
        print(json.dumps(res), file=real_stdout, flush=True)
    return 0


if __name__ == "__main__":
    sys.exit(main())
