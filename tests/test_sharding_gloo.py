"""world_size-2 gloo test (CPU) of the multi-GPU host logic: contiguous frame-range sharding with a
1-frame halo and the gather of per-frame counts -- the only exchange the path has."""
import os
import socket
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_frames, q):
    from slam_cin0051_b200.sharding import frame_range, gather_counts
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    lo, hi, halo = frame_range(n_frames, rank, world)
    # a stand-in for the per-frame device counts: a known function of the global frame index
    local = np.stack([np.arange(lo, hi) * 3 + 1, np.arange(lo, hi) * 5 + 2, np.arange(lo, hi)], 1).astype(np.int32)
    full = gather_counts(local, n_frames, rank, world)
    q.put((rank, lo, hi, halo, full))
    dist.destroy_process_group()


def test_frame_ranges_cover_and_halo():
    from slam_cin0051_b200.sharding import frame_range
    for n in (1, 2, 7, 8, 1000):
        for world in (1, 2, 3, 8):
            seen = []
            for r in range(world):
                lo, hi, halo = frame_range(n, r, world)
                seen.extend(range(lo, hi))
                assert halo == (1 if hi < n and hi > lo else 0)  # the next frame is re-extracted locally, no exchange
            assert seen == list(range(n))


def test_gather_counts_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    n = 37
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = np.stack([np.arange(n) * 3 + 1, np.arange(n) * 5 + 2, np.arange(n)], 1).astype(np.int32)
    for rank, lo, hi, halo, full in res:
        assert np.array_equal(full, want)
