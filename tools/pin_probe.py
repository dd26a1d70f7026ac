"""H2D bandwidth from (a) cudaHostAlloc (torch pin_memory) and (b) an anonymous mmap advised MADV_HUGEPAGE, touched, then cudaHostRegister'ed."""
import ctypes, mmap, sys, time
import numpy as np
import torch

N = 467 * 1000 * 1000
dev = torch.empty(N, dtype=torch.uint8, device="cuda")

def bw(host_t, tag):
    s = torch.cuda.Stream()
    for rep in range(3):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(s):
            e0.record()
            for _ in range(4):
                dev.copy_(host_t, non_blocking=True)
            e1.record()
        torch.cuda.synchronize()
        print(tag, rep, round(4 * N / (e0.elapsed_time(e1) * 1e-3) / 1e9, 1), "GB/s", flush=True)

print(open("/sys/kernel/mm/transparent_hugepage/enabled").read().strip(), "|", open("/sys/kernel/mm/transparent_hugepage/defrag").read().strip())
order = sys.argv[1] if len(sys.argv) > 1 else "ab"
for which in order:
    if which == "a":
        t0 = time.time()
        h = torch.empty(N, dtype=torch.uint8, pin_memory=True)
        h.fill_(1)
        print("cudaHostAlloc + fill", round(time.time() - t0, 2), "s")
        bw(h, "hostalloc")
        del h
    else:
        t0 = time.time()
        size = (N + (2 << 20) - 1) // (2 << 20) * (2 << 20) + (2 << 20)
        mm = mmap.mmap(-1, size, flags=mmap.MAP_PRIVATE | mmap.MAP_ANONYMOUS)
        buf = (ctypes.c_char * size).from_buffer(mm)
        addr = ctypes.addressof(buf)
        aligned = (addr + (2 << 20) - 1) // (2 << 20) * (2 << 20)
        libc = ctypes.CDLL(None, use_errno=True)
        rc = libc.madvise(ctypes.c_void_p(aligned), ctypes.c_size_t(size - (2 << 20)), 14)  # MADV_HUGEPAGE
        arr = np.frombuffer(mm, dtype=np.uint8, count=N, offset=aligned - addr)
        arr[:] = 1
        t = torch.from_numpy(arr)
        r = torch.cuda.cudart().cudaHostRegister(t.data_ptr(), N, 0)
        print("mmap+THP madvise rc", rc, "register", r, round(time.time() - t0, 2), "s")
        try:
            print("AnonHugePages:", [l for l in open("/proc/self/smaps_rollup") if "AnonHuge" in l])
        except OSError:
            pass
        bw(t, "thp-registered")
