"""Edge cases run against the -DSLAMCU_DEBUG_BOUNDS build of the library (SLAMCU_LIB=.../libslamcu_dbg.so): every list / tile /
gather index the kernels compute is checked on the device and traps when out of range (a trap surfaces as a CUDA error on the
next call).  Driven by tests/test_gpu_debug_bounds.py; prints `ok <n cases>` when nothing trapped."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import slam_cin0051_b200 as S  # noqa: E402
from slam_cin0051_b200 import _lib  # noqa: E402
from slam_cin0051_b200.synth import make_sequence  # noqa: E402

assert _lib.LIB_PATH.endswith("libslamcu_dbg.so"), _lib.LIB_PATH
DATA = os.path.join(ROOT, "test", "data")
ctx = S.Context(0)
rng = np.random.default_rng(0)
n = 0
ORB = dict(IntensityThreshold=20, ContiguousPixelsThreshold=9, NonMaxSuppression=1, SuppressionWindowSize=3, PatchSize=31, NumBRIEFPairs=256)
for mode in ("reference", "orb"):
    for shape in ((7, 9), (8, 8), (31, 33), (40, 40), (63, 65), (97, 131), (129, 255), (240, 320), (376, 1241)):
        for kind in ("noise", "flat", "scene"):
            img = (rng.integers(0, 256, shape, dtype=np.uint8) if kind == "noise" else np.full(shape, 90, np.uint8) if kind == "flat"
                   else make_sequence(shape[0], shape[1], 1, 14, seed=shape[0])[0])
            for extra in ((dict(NumLevels=8, ScaleFactor=1.2, MaxFeatures=2000, FastThreshold=5), dict(NumLevels=3, ScaleFactor=2.0, MaxFeatures=50)) if mode == "orb"
                          else (dict(), dict(IntensityThreshold=5, ContiguousPixelsThreshold=0, SuppressionWindowSize=3))):
                det = S.FeatureDetector({**ORB, **extra} if mode == "orb" else {**ORB, "ContiguousPixelsThreshold": 12, "SuppressionWindowSize": 12, **extra}, ctx)
                try:
                    k, d = det.detect_and_compute(img)
                except RuntimeError as e:  # images too small for a level: a clean error, not a trap
                    assert "too small" in str(e), e
                    continue
                ctx.synchronize()
                n += 1
# sequences on noise: the per-level candidate lists overflow (status bits), nothing may be written past their ends
frames = rng.integers(0, 256, (6, 120, 200), dtype=np.uint8)
for cfg, mcfg in (("feature_detector_orb.yml", "feature_matcher_orb.yml"), ("feature_detector.yml", "feature_matcher.yml")):
    det = S.FeatureDetector(os.path.join(DATA, cfg), ctx)
    mat = S.FeatureMatcher(os.path.join(DATA, mcfg), ctx)
    for cap in (64, 512, 4096):
        seq = S.FrameSequence(120, 200, 6, desc_bytes=32, max_raw_corners=256, max_keypoints=cap, context=ctx)
        seq.upload(frames)
        seq.extract(det)
        seq.match_consecutive(mat, with_keypoints="orb" not in cfg)
        c = seq.counts()
        assert (c[:, 0] <= cap).all()
        n += 1
# matcher: ragged sizes, tiny sets, every slice count
mat = S.FeatureMatcher(dict(DistanceType="HAMMING", FilterMatches=1, GoodMatchesCount=20, UseRatioTest=1, RatioTestThreshold=0.9), ctx)
for n1, n2, w in ((1, 1, 32), (2, 1, 32), (255, 257, 32), (1000, 3, 32), (3, 1000, 32), (129, 4097, 32), (64, 64, 5), (300, 300, 64)):
    d1 = rng.integers(0, 256, (n1, w), dtype=np.uint8)
    d2 = rng.integers(0, 256, (n2, w), dtype=np.uint8)
    for slices in (0, 1, 7, 32):
        mat.set_train_slices(slices)
        mat.match(d1, d2)
        mat.knn2(d1, d2)
        n += 1
ctx.synchronize()
print("ok", n)
