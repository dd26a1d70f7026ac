"""Quick diagnostic run on the GPU box (not a pytest file): prints where parity first breaks."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, cv2
import slam_cin0051_b200 as s
from oracle import ref_oracle as ro

D = os.path.join(ROOT, "test", "data")
ctx = s.Context.default()
det = s.FeatureDetector(os.path.join(D, "feature_detector.yml"), ctx)
mat = s.FeatureMatcher(os.path.join(D, "feature_matcher.yml"), ctx)
print("pattern ok", np.array_equal(det.brief_pattern, ro.brief_pattern()), "weights ok", np.array_equal(det.blur_weights, ro.blur_weights()))
im0 = cv2.imread(os.path.join(D, "images/0000000000.png"), 0)
im1 = cv2.imread(os.path.join(D, "images/0000000001.png"), 0)
def bits(a): return np.ascontiguousarray(a, np.float32).view(np.uint32)
for name, im in [("kitti0", im0), ("kitti1", im1)]:
    raw_g = det.fast_corners(im); raw_w = ro.fast_scan(im, scored=True)
    print(name, "raw", len(raw_g), len(raw_w), "equal", len(raw_g) == len(raw_w) and all(np.array_equal(raw_g[f], raw_w[f]) for f in ("x", "y", "response")))
    kg = det.detect(im); kw = ro.detect(im)
    same = len(kg) == len(kw) and all(np.array_equal(kg[f], kw[f]) for f in ("x", "y", "response"))
    print(name, "detect", len(kg), len(kw), "equal", same)
    if not same:
        m = min(len(kg), len(kw)); bad = np.nonzero((kg["x"][:m] != kw["x"][:m]) | (kg["y"][:m] != kw["y"][:m]))[0]
        print("  first diffs", bad[:10], kg[:3], kw[:3])
    bg = det.gaussian_blur(im); bw = ro.gaussian_blur(im)
    print(name, "blur mismatches", int((bg != bw).sum()))
    t = time.time(); gk, gd = det.detect_and_compute(im); dt = time.time() - t
    wk, wd = ro.detect_and_compute(im)
    okk = len(gk) == len(wk)
    print(name, "dac n", len(gk), len(wk), "angle bit-mismatch", int((bits(gk["angle"]) != bits(wk["angle"])).sum()) if okk else "n/a",
          "desc row mismatch", int((gd != wd).any(1).sum()) if okk else "n/a", f"gpu {dt*1e3:.2f} ms")
k0, d0 = ro.detect_and_compute(im0); k1, d1 = ro.detect_and_compute(im1)
for kp in (True, False):
    g = mat.match(d0, d1, k0 if kp else None, k1 if kp else None)
    q, t_, d = ro.match(d0, d1, k0 if kp else None, k1 if kp else None, stage=1)
    print("match kp", kp, len(g), len(q), "equal", len(g) == len(q) and np.array_equal(g["queryIdx"], q) and np.array_equal(g["trainIdx"], t_) and np.array_equal(g["distance"], d))
for i in range(3):
    t = time.time(); det.detect_and_compute(im0); print(f"dac {1e3*(time.time()-t):.2f} ms", end="  ")
    t = time.time(); mat.match(d0, d1, k0, k1); print(f"match {1e3*(time.time()-t):.2f} ms")
print("launches", ctx.launch_count)
