"""GPU-box probe: host<->device copy rates relevant to the end-to-end leg (not part of the product).

  python tools/xfer_probe.py                      single GPU: copy shapes (linear vs strided) and H2D || D2H overlap
  torchrun --nproc-per-node N tools/xfer_probe.py --ranks
                                                 every rank moves bench.py's per-step bytes (467 MB H2D + 117 MB D2H, pinned,
                                                 concurrently on two streams, no kernels) at the same time: the aggregate is the
                                                 HOST-LINK CEILING the end-to-end leg can reach at N GPUs.  One JSON line from rank 0.
"""
import json
import os
import sys
import time

import torch


def t(fn, n=5):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n


def ranks_mode():
    import torch.distributed as dist
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    h2d_b, d2h_b = 1001 * 376 * 1241, 116_712_460
    hin = torch.empty(h2d_b, dtype=torch.uint8, pin_memory=True); hin.fill_(1)
    din = torch.empty(h2d_b, dtype=torch.uint8, device="cuda")
    hout = torch.empty(d2h_b, dtype=torch.uint8, pin_memory=True); hout.fill_(1)
    dout = torch.empty(d2h_b, dtype=torch.uint8, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

    def both():
        with torch.cuda.stream(s1): din.copy_(hin, non_blocking=True)
        with torch.cuda.stream(s2): hout.copy_(dout, non_blocking=True)

    def h2d_only():
        with torch.cuda.stream(s1): din.copy_(hin, non_blocking=True)
    res = {}
    for name, fn in (("h2d_and_d2h", both), ("h2d_only", h2d_only)):
        fn(); torch.cuda.synchronize()
        if world > 1: dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(10): fn()
        torch.cuda.synchronize()
        dt = torch.tensor([(time.perf_counter() - t0) / 10], dtype=torch.float64, device="cuda")
        if world > 1: dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        res[name] = float(dt.item())
    if rank == 0:
        step = res["h2d_and_d2h"]
        print(json.dumps({"tool": "xfer_probe --ranks", "n_gpus": world, "h2d_bytes": h2d_b, "d2h_bytes": d2h_b, "ms_per_step_copies_only": step * 1e3,
                          "aggregate_gbs": world * (h2d_b + d2h_b) / step / 1e9, "per_rank_gbs": (h2d_b + d2h_b) / step / 1e9,
                          "frames_per_s_ceiling": world * 1000 / step, "h2d_only_ms": res["h2d_only"] * 1e3,
                          "h2d_only_aggregate_gbs": world * h2d_b / res["h2d_only"] / 1e9}), flush=True)
    if world > 1: dist.destroy_process_group()


def single_mode():
    B, R, C, P = 512, 376, 1241, 1280
    h = torch.empty((B, R, C), dtype=torch.uint8, pin_memory=True)
    d1 = torch.empty((B, R, C), dtype=torch.uint8, device="cuda")
    d2 = torch.empty((B, R, P), dtype=torch.uint8, device="cuda")
    s = t(lambda: d1.copy_(h, non_blocking=True)); print(f"H2D 1D   {h.numel()/s/1e9:.1f} GB/s  {s*1e3:.2f} ms")
    s = t(lambda: d2[:, :, :C].copy_(h, non_blocking=True)); print(f"H2D 2D(torch) {h.numel()/s/1e9:.1f} GB/s  {s*1e3:.2f} ms")
    import ctypes
    rt = ctypes.CDLL("libcudart.so.12")
    rt.cudaMemcpy2DAsync.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_size_t, ctypes.c_size_t, ctypes.c_int, ctypes.c_void_p]

    def m2d(dst, dp, src, sp, w, hh, kind):
        r = rt.cudaMemcpy2DAsync(dst, dp, src, sp, w, hh, kind, None); assert r == 0, r
    s = t(lambda: m2d(d2.data_ptr(), P, h.data_ptr(), C, C, B * R, 1)); print(f"H2D cudaMemcpy2D w=1241 {h.numel()/s/1e9:.1f} GB/s  {s*1e3:.2f} ms")
    hk = torch.empty((B, 2560, 32), dtype=torch.uint8, pin_memory=True)
    dk = torch.empty((B, 2560, 32), dtype=torch.uint8, device="cuda")
    s = t(lambda: hk.copy_(dk, non_blocking=True)); print(f"D2H 1D 42MB  {hk.numel()/s/1e9:.1f} GB/s  {s*1e3:.2f} ms")
    s = t(lambda: m2d(hk.data_ptr(), 32, dk.data_ptr(), 32, 32, B * 2560, 2)); print(f"D2H cudaMemcpy2D w=32 rows=1.3M {hk.numel()/s/1e9:.1f} GB/s  {s*1e3:.2f} ms")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

    def both():
        with torch.cuda.stream(s1): d1.copy_(h, non_blocking=True)
        with torch.cuda.stream(s2): hk.copy_(dk, non_blocking=True)
    s = t(both); print(f"H2D 239MB || D2H 42MB: {s*1e3:.2f} ms")


if __name__ == "__main__":
    ranks_mode() if "--ranks" in sys.argv else single_mode()
