"""Frame-range sharding of a sequence over ranks (one process per GPU).

Frames are independent for extraction and frame pairs are independent for matching, so rank r owns the
contiguous range [lo, hi) and additionally extracts frame `hi` (a 1-frame halo that is recomputed, not
exchanged) so every pair (f, f+1) with f in [lo, hi) is local.  The only collective is the gather of the
per-frame int32 count triplets {n_keypoints, n_matches, n_inliers} (12 B / frame) over NCCL (gloo on CPU).
"""
from __future__ import annotations

import numpy as np


def frame_range(n_frames: int, rank: int, world: int):
    """Returns (lo, hi, halo): frames [lo, hi) are owned; halo = 1 if frame `hi` must also be extracted."""
    base, rem = divmod(n_frames, world)
    lo = rank * base + min(rank, rem)
    hi = lo + base + (1 if rank < rem else 0)
    halo = 1 if (hi < n_frames and hi > lo) else 0
    return lo, hi, halo


def gather_counts(local_counts: np.ndarray, n_frames: int, rank: int, world: int, device=None) -> np.ndarray:
    """all_gather of per-frame count rows; every rank returns the full (n_frames, k) table."""
    import torch
    import torch.distributed as dist

    local_counts = np.ascontiguousarray(local_counts, np.int32)
    k = local_counts.shape[1]
    if world == 1:
        return local_counts.copy()
    cap = -(-n_frames // world)
    buf = torch.zeros((cap, k), dtype=torch.int32, device=device)
    buf[: len(local_counts)] = torch.from_numpy(local_counts).to(buf.device)
    out = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(out, buf)
    rows = []
    for r in range(world):
        lo, hi, _ = frame_range(n_frames, r, world)
        rows.append(out[r][: hi - lo].cpu().numpy())
    return np.concatenate(rows, 0)
