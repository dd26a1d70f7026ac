"""per-pair match counts: resident one-lane vs pipelined chunks (two lanes), repeated"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import bench
import slam_cin0051_b200 as S
w = bench.WORKLOADS["kitti"]
NF = 1000
frames = bench.make_frames(w, 0, NF)
host = torch.from_numpy(np.ascontiguousarray(np.stack(frames))).pin_memory()
ctx = S.Context(0)
data = os.path.join(bench.ROOT, "test", "data")
det = S.FeatureDetector(os.path.join(data, "feature_detector_orb.yml"), ctx)
mat = S.FeatureMatcher(os.path.join(data, "feature_matcher_orb.yml"), ctx)
seq = S.FrameSequence(w["rows"], w["cols"], NF, desc_bytes=det.descriptor_bytes, max_keypoints=w["max_kp"], context=ctx)
seq.upload_ptr(host.data_ptr(), NF)
seq.extract_match(det, mat, 0, NF, with_keypoints=False, chunk=0)
ctx.synchronize()
base = seq.counts().copy()
print("resident matches", base[:NF - 1, 1].sum())
cnt = torch.empty((NF, 4), dtype=torch.int32).pin_memory()
cap = NF * w["max_kp"]
k = torch.empty((cap, 5), dtype=torch.float32).pin_memory(); d = torch.empty((cap, 32), dtype=torch.uint8).pin_memory(); m = torch.empty((cap, 3), dtype=torch.int32).pin_memory()
for chunk in (125, 125, 100, 50):
    seq.process_dense_ptrs(det, mat, host.data_ptr(), NF, chunk=chunk, with_keypoints=False, kps_ptr=k.data_ptr(), desc_ptr=d.data_ptr(), matches_ptr=m.data_ptr(),
                           counts_ptr=cnt.data_ptr(), kp_capacity=cap, match_capacity=cap)
    seq.wait()
    c = cnt.numpy()
    diff = np.nonzero(c[:NF - 1, 1] != base[:NF - 1, 1])[0]
    print("chunk", chunk, "matches", c[:NF - 1, 1].sum(), "n differing", len(diff), "pairs", diff.tolist())
