"""Row-band sharding of the depth filter across the GPUs of one node.

One process per GPU (torch.distributed, NCCL over NVLink/NVSwitch).  Pixels are independent
(dense_mapping/test_monocular_mapping.cpp:366,546-564 touch only their own map entry), so the
state maps are split into contiguous interior row bands, one per rank; every rank needs the whole
current frame (an epipolar segment can reach anywhere inside the border).  Per frame, rank 0
broadcasts the u8 frame; the pose table of the sequence is broadcast once; at the end the bands
are gathered on rank 0.  There is no exchange inside a frame.

The reference has no distributed code at all; this module is the new piece.
torch is used for device buffers, streams and the collectives only — the update itself is the
C-ABI call dmf_update_device() on each rank's context.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import numpy as np

from .depth_filter import DepthFilter
from .se3 import SE3


def band_rows(height: int, border: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous interior-row band [r0, r1) of `rank`; bands tile [border, height-border) exactly.
    Rank 0 / world-1 additionally own the top / bottom border rows so that the gathered maps cover
    the full image."""
    interior = height - 2 * border
    base, rem = divmod(interior, world)
    start = border + rank * base + min(rank, rem)
    stop = start + base + (1 if rank < rem else 0)
    if rank == 0:
        start = 0
    if rank == world - 1:
        stop = height
    return start, stop


def cyclic_rows(height: int, border: int, block_rows: int, world: int, rank: int) -> np.ndarray:
    """Interior rows of `rank` under block-cyclic ownership (same rule as dmf_create_cyclic)."""
    lo, hi = border, height - border
    rows = []
    k = 0
    n_blocks = -(-(hi - lo) // block_rows)
    rev_round = n_blocks // world if n_blocks % world else -1  # an incomplete last round is dealt from the highest rank down
    while True:  # round k deals blocks k*world .. k*world + world-1, odd rounds in reverse (boustrophedon)
        pos = (world - 1 - rank) if ((k & 1) or k == rev_round) else rank
        y0 = lo + (k * world + pos) * block_rows
        if y0 >= hi:
            break
        rows.extend(range(y0, min(y0 + block_rows, hi)))
        k += 1
    return np.array(rows, dtype=np.int64)


class _DevArray:
    """Expose a raw device pointer to torch through __cuda_array_interface__."""

    def __init__(self, ptr: int, shape: Tuple[int, ...], typestr: str):
        self.__cuda_array_interface__ = {"shape": shape, "typestr": typestr, "data": (ptr, False), "version": 3,
                                         "strides": None}


class ShardedDepthFilter:
    """Depth filter whose state is split into row bands over the ranks of a process group."""

    def __init__(self, params, *, group=None, device: Optional[int] = None, n_ring: int = 4,
                 layout: str = "cyclic", block_rows: int = 8, transport: str = "auto"):
        """layout "cyclic" (default): blocks of `block_rows` interior rows dealt round-robin to the ranks —
        convergence varies smoothly down the image, so this balances the per-rank work; "bands": one
        contiguous band per rank (SURVEY.md 8e first choice; measured 4x imbalance on the 4K sequence).
        transport "ring" (default on CUDA with more than one rank): rank 0 publishes every frame into a ring in its
        HBM and every rank pulls it with a copy engine over NVLink (frame_ring.py: no SM, no collective kernel);
        "broadcast": one NCCL / gloo broadcast per frame on a side stream (the CPU protocol test, and the fallback
        where stream memory operations are unavailable)."""
        import torch
        import torch.distributed as dist

        self.torch, self.dist = torch, dist
        self.group = group
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.params = params
        self.layout = layout if self.world > 1 else "bands"
        self.block_rows = block_rows
        r0, r1 = band_rows(params.height, params.border, self.world, self.rank)
        self.rows = (r0, r1)
        self.pitch = (params.width + 15) // 16 * 16
        self.H, self.W = params.height, params.width
        if n_ring < 2:
            raise ValueError("n_ring must be >= 2: the one-frame look-ahead of prefetch() needs a second ring slot")
        self._k = 0
        self._ring_events = [None] * n_ring
        self._queue = []  # frames announced with prefetch() / prefetch_host(), oldest first: (buffer, comm stream, ring slot)
        self.n_ring = n_ring
        self.transport = transport
        self.frame_ring = self.frame_ring_out = None
        self.frame_rings = []          # ring p (producer = rank p) opened as consumer `rank`; ring 0 = frame_ring
        self.shared_frames = None
        self._published = self._consumed = 0
        self._shared_pub = self._shared_con = 0
        self._attach(device, n_ring)
        if self.transport == "auto":
            self.transport = "ring" if (self.world > 1 and getattr(self.tdev, "type", "cpu") == "cuda" and self.filter is not None) else "broadcast"
        if self.transport == "ring":
            self._attach_ring()

    # The three hooks below are the only places that touch CUDA; tests/test_sharded_gloo.py
    # overrides them with an oracle-backed band on CPU tensors to exercise the protocol
    # (band partition, frame / pose broadcast, gather) under gloo with world_size 2.
    def _attach(self, device, n_ring) -> None:
        torch = self.torch
        self.device = torch.cuda.current_device() if device is None else device
        if self.layout == "cyclic":
            self.filter = DepthFilter(self.params, device=self.device, cyclic=(self.block_rows, self.world, self.rank))
        else:
            self.filter = DepthFilter(self.params, device=self.device, rows=self.rows)
        dev = torch.device("cuda", self.device)
        self.tdev = dev
        self.ring = [torch.empty((self.H, self.pitch), dtype=torch.uint8, device=dev) for _ in range(n_ring)]
        self.comm_stream = torch.cuda.Stream(device=dev)
        # one side stream per ring slot: an update waits for ITS frame only, not for frames announced later
        self.comm_streams = [self.comm_stream] + [torch.cuda.Stream(device=dev) for _ in range(n_ring - 1)]
        self.ctx_stream = torch.cuda.ExternalStream(self.filter.stream(), device=dev)
        d_ptr, c_ptr, pitch = self.filter.device_state()
        assert pitch == self.W * 8
        self.depth_t = torch.as_tensor(_DevArray(d_ptr, (self.H, self.W), "<f8"), device=dev)
        self.cov2_t = torch.as_tensor(_DevArray(c_ptr, (self.H, self.W), "<f8"), device=dev)

    def _attach_ring(self) -> None:
        """Every rank creates a frame ring in its own HBM and opens all of them as consumer number `rank`.  Ring 0 (rank 0)
        carries the frames that rank 0 alone holds (device-resident sequences, host frames private to rank 0); with host
        frames in shared memory (shared_host_frames) frame k is published by rank k mod N into ring k mod N, so that the
        uploads use every GPU's PCIe link."""
        from .frame_ring import FrameRing
        dist = self.dist
        self.frame_ring_out = FrameRing.create(self.device, self.n_ring, self.W, self.H, self.world)
        handles = [None] * self.world
        if self.world > 1:
            dist.all_gather_object(handles, self.frame_ring_out.handle, group=self.group)
        else:
            handles[0] = self.frame_ring_out.handle
        self.frame_rings = [FrameRing.open(self.device, h, self.rank) for h in handles]
        self.frame_ring = self.frame_rings[0]
        if self.world > 1:
            dist.barrier(group=self.group)

    # -- host frames shared by all ranks of the node -------------------------------------------------------
    def shared_host_frames(self, n_frames: int):
        """Collective.  A (n_frames, H, W) uint8 numpy array in POSIX shared memory, page-locked in every rank; rank 0
        fills it (then call a barrier).  prefetch_shared(i) / update_shared(pose) then publish frame i from rank
        i mod N: every rank uploads 1/N of the frames over its own PCIe link."""
        import os

        from .frame_ring import SharedHostFrames
        if self.transport != "ring":
            raise RuntimeError("shared host frames need the ring transport")
        box = [f"dmf_frames_{os.getpid()}_{id(self) & 0xffffff:x}" if self.rank == 0 else None]
        if self.rank == 0:
            self.shared_frames = SharedHostFrames(box[0], (n_frames, self.H, self.W), create=True)
        if self.world > 1:
            self.dist.broadcast_object_list(box, src=0, group=self.group)
            if self.rank != 0:
                self.shared_frames = SharedHostFrames(box[0], (n_frames, self.H, self.W), create=False)
            self.dist.barrier(group=self.group)
        self._shared_pub = self._shared_con = 0
        return self.shared_frames.array

    def prefetch_shared(self, i: int) -> None:
        """Announce frame i of the shared host frames (call in the same order on every rank): its producer, rank
        (n-th announced frame) mod N, enqueues the H2D copy into its ring."""
        k = self._shared_pub
        self._shared_pub += 1
        if k - self._shared_con >= self.n_ring * self.world:
            raise RuntimeError("more frames announced than ring slots: call update_shared() before announcing more")
        if k % self.world == self.rank:
            self.frame_ring_out.publish(self.shared_frames.frame_ptr(i), self.W, None)

    def update_shared(self, pose) -> None:
        """update() against the next announced frame of the shared host frames."""
        k = self._shared_con
        self._shared_con += 1
        self.filter.update_ring(self.frame_rings[k % self.world], pose)

    def _ring_publish(self, frame_dev, host_frame) -> None:
        if self.rank != 0:
            return
        if self._published - self._consumed >= self.n_ring:
            raise RuntimeError("more frames announced than ring slots: call update() before announcing the next frame")
        if host_frame is not None:  # pinned host frame: H2D by a copy engine straight into the ring slot
            self.frame_ring_out.publish(host_frame.data_ptr(), host_frame.stride(0), None)
        else:
            self.frame_ring_out.publish(frame_dev.data_ptr(), frame_dev.stride(0), self.torch.cuda.current_stream(self.tdev).cuda_stream)
        self._published += 1

    def _ring_update(self, frame_dev, host_frame, pose) -> None:
        if self.rank == 0 and self._published == self._consumed:
            self._ring_publish(frame_dev, host_frame)
        self.filter.update_ring(self.frame_ring, pose)
        self._consumed += 1

    def _set_reference(self, buf) -> None:
        self.filter.set_reference_device(buf.data_ptr(), self.pitch)
        self.filter.sync()

    def _launch(self, buf, pose, after_comm) -> None:
        """after_comm: the CUDA stream whose queued work produces `buf`; None: torch's current stream (a frame rendered
        or copied by the caller just before this call is complete once that stream reaches this point)."""
        if after_comm is None or after_comm is False:
            after_comm = self.torch.cuda.current_stream(self.tdev)
        self.filter.update_device(buf.data_ptr(), self.pitch, pose, wait_stream=after_comm.cuda_stream)

    # -- setup ---------------------------------------------------------------------------
    def set_reference(self, ref_dev) -> None:
        """ref_dev: torch uint8 (H, pitch) tensor valid on rank 0; broadcast to all ranks."""
        torch, dist = self.torch, self.dist
        buf = ref_dev if self.rank == 0 else torch.empty((self.H, self.pitch), dtype=torch.uint8, device=self.tdev)
        if self.world > 1:
            dist.broadcast(buf, src=0, group=self.group)
            self._sync_current()
        self._set_reference(buf)

    def _sync_current(self) -> None:
        if self.tdev.type == "cuda":
            self.torch.cuda.current_stream().synchronize()

    def fill_state(self, d0: float = 3.0, c0: float = 3.0) -> None:
        self.filter.fill_state(d0, c0)

    def broadcast_poses(self, poses: Optional[Sequence[SE3]]) -> List[Tuple[tuple, tuple]]:
        """Rank 0 passes the T_C_R list of the sequence; every rank gets it back (one collective)."""
        torch, dist = self.torch, self.dist
        if self.world == 1:
            return [(T.q, T.t) for T in poses]
        n = torch.tensor([len(poses) if self.rank == 0 else 0], dtype=torch.int64, device=self.tdev)
        dist.broadcast(n, src=0, group=self.group)
        tab = torch.empty((int(n.item()), 7), dtype=torch.float64, device=self.tdev)
        if self.rank == 0:
            tab.copy_(torch.tensor([list(T.q) + list(T.t) for T in poses], dtype=torch.float64))
        dist.broadcast(tab, src=0, group=self.group)
        host = tab.cpu().numpy()
        return [(tuple(r[:4]), tuple(r[4:])) for r in host]

    # -- per frame -------------------------------------------------------------------------
    def _move_frame(self, frame_dev, host_frame) -> None:
        """Enqueue the transfer of one frame to every rank (rank 0: optional H2D copy, then NCCL broadcast on a side
        stream) into the next ring slot and remember it for the update that will consume it."""
        torch, dist = self.torch, self.dist
        b = self._k % len(self.ring)
        self._k += 1
        buf = self.ring[b] if (self.rank != 0 or host_frame is not None) else frame_dev
        if self.tdev.type != "cuda":  # gloo / CPU protocol test: synchronous
            if self.rank == 0 and host_frame is not None:
                buf[:, : self.W].copy_(host_frame)
            if self.world > 1:
                dist.broadcast(buf, src=0, group=self.group)
            self._queue.append((buf, None, b))
            return
        cs = self.comm_streams[b]
        cs.wait_stream(torch.cuda.current_stream(self.tdev))  # the caller's stream may still be writing the frame
        with torch.cuda.stream(cs):
            if self._ring_events[b] is not None:
                cs.wait_event(self._ring_events[b])  # the kernel that last read ring[b] is done
            if self.rank == 0 and host_frame is not None:
                buf[:, : self.W].copy_(host_frame, non_blocking=True)
            if self.world > 1:
                dist.broadcast(buf, src=0, group=self.group)
        self._queue.append((buf, cs, b))

    def _consume(self, pose) -> None:
        buf, cs, b = self._queue.pop(0)
        self._launch(buf, pose, cs)
        if cs is not None:
            ev = self.torch.cuda.Event()
            ev.record(self.ctx_stream)
            self._ring_events[b] = ev

    def prefetch(self, frame_dev) -> None:
        """Announce the frame of a FUTURE update (frames are consumed in the order they were announced).  Its
        broadcast is enqueued now, so it can run in a gap one update earlier: ncc_kernel keeps every SM busy with
        persistent CTAs, and an NCCL kernel enqueued right before the update it feeds would have to wait for the
        previous ncc_kernel to drain, putting the broadcast and the frame-only precompute on the critical path."""
        if self.transport == "ring":
            self._ring_publish(frame_dev, None)
        elif self.world == 1:
            self._queue.append((frame_dev, None, 0))
        else:
            self._move_frame(frame_dev, None)

    def prefetch_host(self, host_frame) -> None:
        """prefetch() for the end-to-end path: `host_frame` is a pinned torch uint8 (H, W) tensor on rank 0."""
        if self.transport == "ring":
            self._ring_publish(None, host_frame)
            return
        self._move_frame(None, host_frame if self.rank == 0 else None)

    def update(self, frame_dev, pose: Tuple[tuple, tuple]) -> None:
        """frame_dev: torch uint8 (H, pitch) tensor on rank 0 (ignored elsewhere, and ignored everywhere if frames
        were announced with prefetch()).  Asynchronous: the broadcast runs on a side stream, multi-buffered against
        the previous frames' kernels."""
        if self.transport == "ring":
            return self._ring_update(frame_dev, None, pose)
        if not self._queue:
            if self.world == 1:
                self._launch(frame_dev, pose, None)
                return
            self._move_frame(frame_dev, None)
        self._consume(pose)

    def update_host(self, host_frame, pose: Tuple[tuple, tuple]) -> None:
        """End-to-end form of update(): `host_frame` is a pinned torch uint8 (H, W) tensor on rank 0 (None
        elsewhere; ignored if frames were announced with prefetch_host()).  Rank 0 copies it to HBM on the side
        stream, the frame is broadcast, every rank updates its rows; all of it overlapped with the previous kernels."""
        if self.transport == "ring":
            return self._ring_update(None, host_frame, pose)
        if not self._queue:
            self._move_frame(None, host_frame if self.rank == 0 else None)
        self._consume(pose)

    # -- results ---------------------------------------------------------------------------
    def gather_state(self):
        """Gather the bands on rank 0: returns (depth, cov2) torch tensors (H, W) on rank 0, else None."""
        torch, dist = self.torch, self.dist
        self._sync_filter()
        if self.world == 1:
            return self.depth_t, self.cov2_t
        send_rows = self._rows_of(self.rank)
        max_rows = max(len(self._rows_of(r)) for r in range(self.world))
        send = torch.zeros((2, max_rows, self.W), dtype=torch.float64, device=self.depth_t.device)
        idx = torch.as_tensor(send_rows, device=self.depth_t.device)
        send[0, : len(send_rows)] = self.depth_t.index_select(0, idx)
        send[1, : len(send_rows)] = self.cov2_t.index_select(0, idx)
        if self.rank == 0:
            recv = [torch.empty_like(send) for _ in range(self.world)]
            dist.gather(send, recv, dst=0, group=self.group)
            for r in range(1, self.world):
                rr = self._rows_of(r)
                ridx = torch.as_tensor(rr, device=self.depth_t.device)
                self.depth_t.index_copy_(0, ridx, recv[r][0, : len(rr)])
                self.cov2_t.index_copy_(0, ridx, recv[r][1, : len(rr)])
            self._sync_current()
            return self.depth_t, self.cov2_t
        dist.gather(send, None, dst=0, group=self.group)
        self._sync_current()
        return None

    def _rows_of(self, r: int) -> np.ndarray:
        """Rows rank r contributes to the gather (its interior rows; the border rows never change)."""
        if self.layout == "cyclic":
            return cyclic_rows(self.H, self.params.border, self.block_rows, self.world, r)
        a, b = band_rows(self.H, self.params.border, self.world, r)
        return np.arange(max(a, self.params.border), min(b, self.H - self.params.border), dtype=np.int64)

    def _sync_filter(self) -> None:
        self.filter.sync()

    def flush(self) -> None:
        """Enqueue the deferred fusion of the last update on the context stream (asynchronous)."""
        if hasattr(self.filter, "flush"):
            self.filter.flush()

    def _local_counters(self, reset: bool) -> dict:
        return self.filter.counters(reset)

    def counters(self, reset: bool = False) -> dict:
        """Work counters summed over ranks (all ranks get the total)."""
        torch, dist = self.torch, self.dist
        c = self._local_counters(reset)
        if self.world == 1:
            return c
        keys = ["interior", "active", "ncc_evals", "accepted"]
        t = torch.tensor([c[k] for k in keys], dtype=torch.int64, device=self.depth_t.device)
        dist.all_reduce(t, group=self.group)
        out = {k: int(v) for k, v in zip(keys, t.tolist())}
        out["frames"] = c["frames"]
        return out

    def close(self) -> None:
        if self.filter is not None:
            self.filter.sync()
        if self.world > 1 and self.frame_ring is not None:
            self.dist.barrier(group=self.group)  # nobody unmaps the ring while a peer is still pulling from it
        for r in list(self.frame_rings) + [self.frame_ring_out]:
            if r is not None:
                r.close()
        self.frame_rings, self.frame_ring, self.frame_ring_out = [], None, None
        if self.shared_frames is not None:
            self.shared_frames.close()
            self.shared_frames = None
        if self.filter is not None:
            self.filter.close()
