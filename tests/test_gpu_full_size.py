"""Parity at BASELINE.json's full sizes (configs 2-5), where the oracles are too slow for every element: live cv2 for
whole-frame results where it finishes in seconds, size-independent properties elsewhere."""
import os

import numpy as np
import pytest

from conftest import DATA
from test_orb_oracle import assert_orb_equal

pytestmark = pytest.mark.gpu

ORB_CFG = dict(IntensityThreshold=20, ContiguousPixelsThreshold=9, NonMaxSuppression=1, SuppressionWindowSize=3,
               PatchSize=31, NumBRIEFPairs=256, NumLevels=8, ScaleFactor=1.2, MaxFeatures=2000)


def test_config4_4k_10000_keypoints_equals_cv2(gpu_ctx):
    pytest.importorskip("cv2")
    import slam_cin0051_b200 as s
    from slam_cin0051_b200.synth import make_sequence
    from tools_golden import orb_canonical
    det = s.FeatureDetector({**ORB_CFG, "MaxFeatures": 10000}, gpu_ctx)
    img = make_sequence(2160, 3840, 1, pitch_px=28, seed=11)[0]
    k, d = det.detect_and_compute(img)
    r = {"x": k["x"], "y": k["y"], "size": k["size"], "angle": k["angle"], "response": k["response"],
         "octave": det.last_octaves(len(k)), "desc": d}
    assert len(k) == 10000
    assert_orb_equal(r, orb_canonical(img, nfeatures=10000))


def test_config3_tum_shape_sequence_with_ransac(gpu_ctx):
    """640x480, 1000 keypoints, consecutive-frame matching + findEssentialMat per pair on a 64-frame sequence:
    every frame equals the single-frame path on sampled frames, every pair's inlier count equals its mask sum, and a
    pure image translation keeps almost all ratio-test matches as RANSAC inliers."""
    import slam_cin0051_b200 as s
    from slam_cin0051_b200.synth import make_sequence
    det = s.FeatureDetector({**ORB_CFG, "MaxFeatures": 1000}, gpu_ctx)
    mat = s.FeatureMatcher(os.path.join(DATA, "feature_matcher_orb.yml"), gpu_ctx)
    n = 64
    frames = make_sequence(480, 640, n, pitch_px=17, seed=3)
    seq = s.FrameSequence(480, 640, n, desc_bytes=32, max_keypoints=1280, context=gpu_ctx)
    seq.upload(frames)
    seq.extract(det)
    seq.match_consecutive(mat, with_keypoints=False)
    seq.essential((525.0, 525.0, 319.5, 239.5))
    counts = seq.counts()
    assert (counts[:, 3] == 0).all() and (counts[:, 0] == 1000).all()
    for f in (0, 17, 63):
        k, d = seq.frame(f)
        k1, d1 = det.detect_and_compute(frames[f])
        assert k.tobytes() == k1.tobytes() and np.array_equal(d, d1)
    for f in range(0, n - 1, 7):
        E, mask, good, iters = seq.essential_result(f)
        assert len(mask) == counts[f, 1] and good == int(mask.sum()) and 1 <= iters <= 1000
        assert good > 0.5 * len(mask)


def test_config5_hamming_properties_at_32k(gpu_ctx):
    """32k x 32k descriptors: matching a set against itself returns every descriptor as its own nearest neighbour at
    distance 0; against a bit-flipped copy the distance equals the number of flipped bits; sampled queries equal numpy."""
    import slam_cin0051_b200 as s
    mat = s.FeatureMatcher(dict(DistanceType="HAMMING", FilterMatches=0, GoodMatchesCount=1, UseRatioTest=0,
                                RatioTestThreshold=1.0), gpu_ctx)
    n = 32768
    rng = np.random.default_rng(7)
    d = rng.integers(0, 256, (n, 32), dtype=np.uint8)
    got = mat.knn2(d, d)
    assert np.array_equal(got["trainIdx0"], np.arange(n)) and (got["distance0"] == 0).all() and (got["distance1"] > 0).all()
    flips = rng.integers(1, 20, n)
    d2 = d.copy()
    for k in range(20):  # flip bit k of byte k in the rows that need more than k flips
        rows = flips > k
        d2[rows, k] ^= np.uint8(1 << (k % 8))
    got = mat.knn2(d2, d)
    assert np.array_equal(got["trainIdx0"], np.arange(n)) and np.array_equal(got["distance0"], flips.astype(np.float32))
    m = mat.match(d2, d)
    assert len(m) == n and np.array_equal(m["queryIdx"], np.arange(n)) and np.array_equal(m["trainIdx"], np.arange(n))
