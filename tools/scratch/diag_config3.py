import os, sys, json, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np, cv2
import slam_cin0051_b200 as s
from oracle import essential_oracle as eo
from slam_cin0051_b200.synth import make_sequence
ctx = s.Context(0)
ORB_CFG = dict(IntensityThreshold=20, ContiguousPixelsThreshold=9, NonMaxSuppression=1, SuppressionWindowSize=3, PatchSize=31, NumBRIEFPairs=256, NumLevels=8, ScaleFactor=1.2, MaxFeatures=1000)
det = s.FeatureDetector(ORB_CFG, ctx)
mat = s.FeatureMatcher(os.path.join(ROOT, "test/data/feature_matcher_orb.yml"), ctx)
n = 1000
frames = np.empty((n, 480, 640), np.uint8)
for g in range(0, n, 16):
    frames[g:g + 16] = make_sequence(480, 640, min(16, n - g), pitch_px=17, seed=g // 16)
K4 = (525.0, 525.0, 319.5, 239.5)
K = np.array([[K4[0], 0, K4[2]], [0, K4[1], K4[3]], [0, 0, 1.0]])
seq = s.FrameSequence(480, 640, n, desc_bytes=32, max_keypoints=1280, context=ctx)
seq.upload(frames); seq.extract(det); seq.match_consecutive(mat, with_keypoints=False); seq.essential(K4)
counts = seq.counts()
def dist(a, b): return float(min(np.abs(a - b).max(), np.abs(a + b).max()))
rows = []
for f in sorted(set(range(0, n - 1, 7)) | set(range(15, n - 1, 112))):
    m = seq.matches(f); ka, _ = seq.frame(f); kb, _ = seq.frame(f + 1)
    E, mask, good, iters = seq.essential_result(f)
    if len(m) < 6: continue
    p1 = np.stack([ka["x"][m["queryIdx"]], ka["y"][m["queryIdx"]]], 1).astype(np.float32)
    p2 = np.stack([kb["x"][m["trainIdx"]], kb["y"][m["trainIdx"]]], 1).astype(np.float32)
    cE, cmask = cv2.findEssentialMat(p1, p2, K, cv2.RANSAC, 0.999, 1.0, 1000)
    cmask = np.zeros(len(m), np.uint8) if cmask is None else cmask.ravel().astype(np.uint8)
    # single-call device path on the same points
    E1, mask1, good1 = s.find_essential(p1, p2, K4, context=ctx)
    r = dict(f=f, n=len(m), iters=int(iters), good=int(good), cv=int(cmask.sum()), eq=bool(np.array_equal(mask, cmask)), single_eq_seq=bool(np.array_equal(mask, mask1)),
             jac=float((mask & cmask).sum() / max(1, (mask | cmask).sum())))
    if good > 0 and cE is not None: r["dE"] = dist(E, np.asarray(cE, np.float64)[:3])
    if not r["eq"] and iters < 200:
        nE, nmask, ngood = eo.find_essential(p1, p2, K4)
        r["np_eq_cv"] = bool(np.array_equal(nmask, cmask)); r["np_eq_dev"] = bool(np.array_equal(nmask, mask)); r["np_good"] = int(ngood)
    rows.append(r)
print(json.dumps(rows))
