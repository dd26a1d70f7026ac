"""GPU-box probe: host<->device copy rates relevant to the end-to-end leg (not part of the product)."""
import time
import torch

def t(fn, n=5):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n

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
hc = torch.empty((B, 4), dtype=torch.int32, pin_memory=True); dc = torch.empty((B,), dtype=torch.int32, device="cuda")
s = t(lambda: m2d(hc.data_ptr(), 16, dc.data_ptr(), 4, 4, B, 2)); print(f"D2H cudaMemcpy2D w=4 rows=512 {s*1e3:.3f} ms")
# overlap: H2D and D2H concurrently on two streams
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def both():
    with torch.cuda.stream(s1): d1.copy_(h, non_blocking=True)
    with torch.cuda.stream(s2): hk.copy_(dk, non_blocking=True)
s = t(both); print(f"H2D 239MB || D2H 42MB: {s*1e3:.2f} ms")
