"""GPU probe: time slamcu_sequence_essential on a TUM-shape synthetic sequence and show the iteration histogram."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import slam_cin0051_b200 as S
from slam_cin0051_b200.synth import make_sequence
from slam_cin0051_b200.sequence import FrameSequence

F = int(sys.argv[1]) if len(sys.argv) > 1 else 512
data = os.path.join(ROOT, "test", "data")
ctx = S.Context(0)
import tempfile
yml = open(os.path.join(data, "feature_detector_orb.yml")).read().replace("MaxFeatures: 2000", "MaxFeatures: 1000")
tmp = os.path.join(data, "_det_probe.yml"); open(tmp, "w").write(yml)
det = S.FeatureDetector(tmp, ctx); os.remove(tmp)
mat = S.FeatureMatcher(os.path.join(data, "feature_matcher_orb.yml"), ctx)
frames = np.concatenate([make_sequence(480, 640, 16, pitch_px=17, seed=1000 + i) for i in range(F // 16)])
seq = FrameSequence(480, 640, F, 32, 0, 1280, ctx)
seq.upload(frames); seq.extract(det); seq.match_consecutive(mat, with_keypoints=False); ctx.synchronize()
K4 = (525.0, 525.0, 319.5, 239.5)
for rep in range(3):
    ctx.profile_enable(True)
    t0 = time.perf_counter(); seq.essential(K4); ctx.synchronize(); t1 = time.perf_counter()
    print("essential wall ms", 1e3 * (t1 - t0), {k: v for k, v in ctx.profile_read().items() if "ess" in k or "pose" in k})
    ctx.profile_enable(False)
its, inl, npt = [], [], []
for p in range(F - 1):
    E, m, ni, nit = seq.essential_result(p)
    its.append(nit); inl.append(ni); npt.append(len(m))
its = np.array(its); inl = np.array(inl); npt = np.array(npt)
print("n_pts mean", npt.mean(), "inlier frac mean", (inl / np.maximum(npt, 1)).mean())
print("iters: min/median/mean/max", its.min(), np.median(its), its.mean(), its.max())
print("iters hist", np.histogram(its, bins=[0, 8, 16, 32, 64, 128, 256, 512, 1001])[0])
# the single-problem goldens (real 3-D scenes with outliers)
import glob
for g in sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "essential_*.npz"))):
    d = np.load(g)
    K = d["K"]; k4 = (K[0, 0], K[1, 1], K[0, 2], K[1, 2])
    S.find_essential(d["p1"], d["p2"], k4, context=ctx)
    t0 = time.perf_counter()
    for _ in range(5): E, mask, good = S.find_essential(d["p1"], d["p2"], k4, context=ctx)
    print(os.path.basename(g), "n", len(d["p1"]), "inliers", good, "ms/call", 1e3 * (time.perf_counter() - t0) / 5)
# the minimal solver alone: one thread per sample
from oracle import essential_oracle as eo
d = np.load(os.path.join(ROOT, "tests", "golden", "essential_kitti01.npz"))
K = d["K"]; k4 = (K[0, 0], K[1, 1], K[0, 2], K[1, 2])
x1, x2 = eo.normalise(d["p1"], k4), eo.normalise(d["p2"], k4)
rng = np.random.default_rng(0)
for S_ in (32, 1024, 16384):
    idx = np.stack([rng.choice(len(x1), 5, replace=False) for _ in range(S_)])
    S.fivept_solve(x1[idx], x2[idx], context=ctx)
    ctx.profile_enable(True) if False else None
    t0 = time.perf_counter(); S.fivept_solve(x1[idx], x2[idx], context=ctx); t1 = time.perf_counter()
    print("fivept_solve", S_, "samples: wall ms (incl. copies)", 1e3 * (t1 - t0))
