import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np, torch
import slam_cin0051_b200 as S
import bench
w = bench.WORKLOADS["kitti"]
B = 1000
ctx = S.Context(0)
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream); ctx.set_stream(stream.cuda_stream)
data = os.path.join(ROOT, "test", "data")
det = S.FeatureDetector(os.path.join(data, "feature_detector_orb.yml"), ctx)
mat = S.FeatureMatcher(os.path.join(data, "feature_matcher_orb.yml"), ctx)
max_kp = 2560
frames = bench.make_frames(w, 0, B)
hf = torch.empty((B, 376, 1241), dtype=torch.uint8, pin_memory=True); hf.numpy()[:] = frames
seqs = [S.FrameSequence(376, 1241, B, desc_bytes=32, max_keypoints=max_kp, context=ctx) for _ in range(2)]
cap = B * max_kp
def outs():
    return (torch.empty((cap, 5), dtype=torch.float32, pin_memory=True), torch.empty((cap, 32), dtype=torch.uint8, pin_memory=True),
            torch.empty((cap, 3), dtype=torch.int32, pin_memory=True), torch.empty((B, 4), dtype=torch.int32, pin_memory=True))
O = [outs(), outs()]
P = [(torch.empty((B, max_kp, 5), dtype=torch.float32, pin_memory=True), torch.empty((B, max_kp, 32), dtype=torch.uint8, pin_memory=True),
      torch.empty((B, max_kp, 3), dtype=torch.int32, pin_memory=True), torch.empty((B, 4), dtype=torch.int32, pin_memory=True)) for _ in range(2)]
def run(dense, chunk, steps=6, what="all"):
    def submit(i):
        k, d, m, c = (O if dense else P)[i % 2]
        kw = dict(kps_ptr=k.data_ptr() if what != "none" else None, desc_ptr=d.data_ptr() if what != "none" else None,
                  matches_ptr=m.data_ptr() if what != "none" else None, counts_ptr=c.data_ptr())
        if dense:
            seqs[i % 2].process_dense_ptrs(det, mat, hf.data_ptr(), B, chunk=chunk, with_keypoints=False, kp_capacity=cap, match_capacity=cap, **kw)
        else:
            seqs[i % 2].process_ptrs(det, mat, hf.data_ptr(), B, chunk=chunk, with_keypoints=False, **kw)
    for i in range(2): submit(i)
    ctx.synchronize(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(steps):
        submit(i)
        if i >= 1: seqs[(i - 1) % 2].wait()
    seqs[(steps - 1) % 2].wait(); ctx.synchronize()
    dt = (time.perf_counter() - t0) / steps
    return dt * 1e3
for chunk in (500, 1000):
    for dense, what in ((False, "all"), (True, "all"), (True, "none"), (False, "none")):
        print(f"chunk {chunk} dense {dense} outputs {what}: {run(dense, chunk, what=what):.2f} ms/step", flush=True)
ctx.profile_enable(True)
run(True, 500, steps=3)
print({k: (round(v[0] / v[1], 3), v[1]) for k, v in ctx.profile_read().items() if k.startswith("dense") or k in ("repitch",)})
ctx.profile_enable(False)
