"""Copy-engine frame ring (include/dmf.h dmf_ring_*, SURVEY.md §8e transport): frames published by one producer are
pulled by every consumer context in order, with stream memory operations as the only synchronisation.  The results must
be bit-identical to a single context fed directly (ref:366,546-564: pixels are independent)."""
import multiprocessing as mp
import os
import time

import numpy as np
import pytest

from slamplay_b200.synth import make_sequence

pytestmark = pytest.mark.gpu

N_FRAMES = 12


def _single_context(seq, frames):
    from slamplay_b200.depth_filter import DepthFilter
    f = DepthFilter(seq.params)
    f.set_reference(frames[0])
    f.fill_state(3.0, 3.0)
    for i in range(1, seq.n_frames):
        f.update(frames[i], seq.T_C_R(i))
    d, c = f.download_state()
    cnt = f.counters()
    f.close()
    return d, c, cnt


@pytest.mark.timeout(180)
def test_ring_two_consumers_one_process_matches_direct_updates():
    import torch
    from slamplay_b200.depth_filter import DepthFilter
    from slamplay_b200.frame_ring import FrameRing
    seq = make_sequence("remode_640x480", n_frames=N_FRAMES)
    p = seq.params
    h, w = seq.shape
    frames = [seq.render_host(i) for i in range(seq.n_frames)]
    d1, c1, cnt1 = _single_context(seq, frames)
    pinned = torch.empty((seq.n_frames, h, w), dtype=torch.uint8, pin_memory=True)
    for i in range(seq.n_frames):
        pinned[i] = torch.from_numpy(frames[i])
    dev = pinned.cuda()
    torch.cuda.synchronize()
    prod = FrameRing.create(0, 3, w, h, 2)  # 3 slots, 11 frames: slots are reused, the release flags matter
    cons = [FrameRing.open(0, prod.handle, k) for k in range(2)]
    parts = [DepthFilter(p, cyclic=(8, 2, k)) for k in range(2)]
    for f in parts:
        f.set_reference(frames[0])
        f.fill_state(3.0, 3.0)
    cur = torch.cuda.current_stream().cuda_stream
    for i in range(1, seq.n_frames):
        if i % 2:   # pinned host frame: H2D by a copy engine into the slot
            prod.publish(pinned[i].data_ptr(), w, None)
        else:       # device frame
            prod.publish(dev[i].data_ptr(), w, cur)
        for f, r in zip(parts, cons):
            f.update_ring(r, seq.T_C_R(i))
    d2, c2 = np.full((h, w), 3.0), np.full((h, w), 3.0)
    tot = {"active": 0, "ncc_evals": 0, "accepted": 0, "interior": 0}
    for f in parts:
        f.download_state(d2, c2)
        for k in tot:
            tot[k] += f.counters()[k]
    assert prod.info()["next_frame"] == seq.n_frames - 1 and cons[1].info()["next_frame"] == seq.n_frames - 1
    for f in parts:
        f.close()
    for r in cons + [prod]:
        r.close()
    assert np.array_equal(d1, d2, equal_nan=True) and np.array_equal(c1, c2, equal_nan=True)
    assert all(tot[k] == cnt1[k] for k in tot)


def _ring_worker(rank, tmp, n_frames):
    """One process per 'rank', both on cuda:0 (the test box has one GPU): rank 0 produces, both consume through IPC."""
    import torch  # noqa: F401  (CUDA context of this process)
    from slamplay_b200.depth_filter import DepthFilter
    from slamplay_b200.frame_ring import FrameRing
    seq = make_sequence("remode_640x480", n_frames=n_frames)
    h, w = seq.shape
    hpath = os.path.join(tmp, "handle.bin")
    prod = None
    if rank == 0:
        prod = FrameRing.create(0, 3, w, h, 2)
        with open(hpath + ".tmp", "wb") as fh:
            fh.write(prod.handle)
        os.rename(hpath + ".tmp", hpath)
    else:
        t0 = time.time()
        while not os.path.exists(hpath):
            if time.time() - t0 > 60:
                raise RuntimeError("no ring handle from rank 0")
            time.sleep(0.05)
    with open(hpath, "rb") as fh:
        handle = fh.read()
    ring = FrameRing.open(0, handle, rank)
    frames = [seq.render_host(i) for i in range(seq.n_frames)]
    f = DepthFilter(seq.params, cyclic=(8, 2, rank))
    f.set_reference(frames[0])
    f.fill_state(3.0, 3.0)
    pinned = torch.empty((seq.n_frames, h, w), dtype=torch.uint8, pin_memory=True)
    for i in range(seq.n_frames):
        pinned[i] = torch.from_numpy(frames[i])
    for i in range(1, seq.n_frames):
        if rank == 0:
            prod.publish(pinned[i].data_ptr(), w, None)
        f.update_ring(ring, seq.T_C_R(i))
        if rank == 1 and i == 4:
            time.sleep(0.5)  # a slow consumer: the producer's stream must wait for the slot, not overwrite it
    d, c = np.full((h, w), np.nan), np.full((h, w), np.nan)
    f.download_state(d, c)
    np.savez(os.path.join(tmp, f"out{rank}.npz"), d=d, c=c, rows=f.owned_rows())
    open(os.path.join(tmp, f"done{rank}"), "w").close()
    # nobody unmaps the ring while the peer may still be pulling from it
    t0 = time.time()
    while not all(os.path.exists(os.path.join(tmp, f"done{k}")) for k in range(2)) and time.time() - t0 < 60:
        time.sleep(0.05)
    f.close()
    ring.close()
    if prod is not None:
        prod.close()


@pytest.mark.timeout(300)
def test_ring_across_two_processes_ipc(tmp_path):
    ctx = mp.get_context("spawn")
    procs = [ctx.Process(target=_ring_worker, args=(r, str(tmp_path), N_FRAMES)) for r in range(2)]
    for pr in procs:
        pr.start()
    deadline = time.time() + 240
    for pr in procs:
        pr.join(max(1.0, deadline - time.time()))
    hung = [pr for pr in procs if pr.is_alive()]
    for pr in hung:
        pr.kill()
    assert not hung, "ring workers did not finish (stream memory operation never satisfied?)"
    assert all(pr.exitcode == 0 for pr in procs), [pr.exitcode for pr in procs]
    seq = make_sequence("remode_640x480", n_frames=N_FRAMES)
    frames = [seq.render_host(i) for i in range(seq.n_frames)]
    d1, c1, _ = _single_context(seq, frames)
    h, w = seq.shape
    d2, c2 = np.full((h, w), 3.0), np.full((h, w), 3.0)
    seen = []
    for r in range(2):
        z = np.load(tmp_path / f"out{r}.npz")
        rows = z["rows"]
        d2[rows], c2[rows] = z["d"][rows], z["c"][rows]
        seen += rows.tolist()
    assert sorted(seen) == list(range(seq.params.border, h - seq.params.border))
    assert np.array_equal(d1, d2, equal_nan=True) and np.array_equal(c1, c2, equal_nan=True)
