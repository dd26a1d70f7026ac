"""GPU parity of the reference-mode path against the CPU oracle.

All comparisons are through the C ABI (via the ctypes mirror classes) and BIT-EXACT: corner
locations and order, integer scores, keypoint order after the std::sort + greedy NMS, blurred bytes,
float angles (compared as bit patterns), descriptor bytes, Hamming distances, match indices and the
std::partial_sort / std::sort output order.
"""
import os

import numpy as np
import pytest

from conftest import DATA

pytestmark = pytest.mark.gpu


def _bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def assert_kps_equal(got, want):
    assert len(got) == len(want)
    for f in ("x", "y", "size", "angle", "response"):
        assert np.array_equal(_bits(got[f]), _bits(want[f])), f


def _images(kitti_pair, tum_pair):
    from slam_cin0051_b200.synth import make_sequence
    syn = make_sequence(376, 1241, 2, pitch_px=14, seed=0)
    return {"kitti0": kitti_pair[0], "kitti1": kitti_pair[1], "tum0": tum_pair[0], "tum1": tum_pair[1],
            "synK0": syn[0], "synK1": syn[1]}


def test_fast_corners_raster_order_and_scores(detector, oracle, kitti_pair, tum_pair):
    for name, img in _images(kitti_pair, tum_pair).items():
        want = oracle.fast_scan(img, scored=True)
        got = detector.fast_corners(img)
        assert_kps_equal(got, want)
    assert len(detector.fast_corners(kitti_pair[0])) == 11329


def test_detect_matches_std_sort_and_greedy_nms(detector, oracle, kitti_pair, tum_pair):
    for name, img in _images(kitti_pair, tum_pair).items():
        assert_kps_equal(detector.detect(img), oracle.detect(img))
    assert len(detector.detect(kitti_pair[0])) == 1145


def test_gaussian_blur_bytes(detector, oracle, kitti_pair, tum_pair):
    imgs = dict(_images(kitti_pair, tum_pair))
    rng = np.random.default_rng(11)
    # white noise and near-flat images: many sums close to k + 0.5, i.e. the exact-path band of the rounding filter
    imgs["noise"] = rng.integers(0, 256, (512, 768), dtype=np.uint8)
    imgs["two_level"] = (rng.integers(0, 2, (300, 500)) * 255).astype(np.uint8)
    imgs["small_range"] = rng.integers(100, 104, (257, 333), dtype=np.uint8)
    for name, img in imgs.items():
        assert np.array_equal(detector.gaussian_blur(img), oracle.gaussian_blur(img)), name


def test_detect_and_compute(detector, oracle, kitti_pair, tum_pair):
    for name, img in _images(kitti_pair, tum_pair).items():
        gk, gd = detector.detect_and_compute(img)
        wk, wd = oracle.detect_and_compute(img)
        assert_kps_equal(gk, wk)
        assert np.array_equal(gd, wd), name
    gk, gd = detector.detect_and_compute(kitti_pair[0])
    assert oracle.fnv1a64(gd.tobytes()) == "f03506a4c32e12dc"


def test_compute_on_given_keypoints(detector, oracle, tum_pair):
    img = tum_pair[0]
    rng = np.random.default_rng(3)
    kps = np.zeros(500, oracle.KP_DTYPE)
    kps["x"] = rng.integers(0, img.shape[1], 500)
    kps["y"] = rng.integers(0, img.shape[0], 500)
    kps["size"] = 6.0
    gk, gd = detector.compute(img, kps)
    wk, wd = oracle.compute(img, kps)
    assert_kps_equal(gk, wk)
    assert np.array_equal(gd, wd)
    ek, ed = detector.compute(img, kps[:0])
    assert len(ek) == 0 and ed.shape == (0, 0)


@pytest.mark.parametrize("cfg", [
    dict(IntensityThreshold=20, ContiguousPixelsThreshold=9, NonMaxSuppression=1, SuppressionWindowSize=7, PatchSize=31, NumBRIEFPairs=256),
    dict(IntensityThreshold=35, ContiguousPixelsThreshold=12, NonMaxSuppression=0, SuppressionWindowSize=12, PatchSize=31, NumBRIEFPairs=256),
    dict(IntensityThreshold=10, ContiguousPixelsThreshold=16, NonMaxSuppression=1, SuppressionWindowSize=20, PatchSize=15, NumBRIEFPairs=64),
    dict(IntensityThreshold=25, ContiguousPixelsThreshold=0, NonMaxSuppression=1, SuppressionWindowSize=12, PatchSize=41, NumBRIEFPairs=512),
    dict(IntensityThreshold=15, ContiguousPixelsThreshold=5, NonMaxSuppression=1, SuppressionWindowSize=3, PatchSize=9, NumBRIEFPairs=8),
])
def test_other_detector_configs(gpu_ctx, oracle, tum_pair, cfg):
    import slam_cin0051_b200 as s
    det = s.FeatureDetector(cfg, gpu_ctx)
    assert np.array_equal(det.brief_pattern, oracle.brief_pattern(cfg["PatchSize"], cfg["NumBRIEFPairs"]))
    img = tum_pair[1]
    gk, gd = det.detect_and_compute(img)
    wk, wd = oracle.detect_and_compute(img, cfg)
    assert_kps_equal(gk, wk)
    assert np.array_equal(gd, wd)


def _assert_matches(got, want):
    q, t, d = want
    assert len(got) == len(q)
    assert np.array_equal(got["queryIdx"], q)
    assert np.array_equal(got["trainIdx"], t)
    assert np.array_equal(_bits(got["distance"]), _bits(d))


def test_match_with_keypoints_and_filter(detector, matcher, oracle, kitti_pair):
    k0, d0 = oracle.detect_and_compute(kitti_pair[0])
    k1, d1 = oracle.detect_and_compute(kitti_pair[1])
    _assert_matches(matcher.match(d0, d1, k0, k1), oracle.match(d0, d1, k0, k1, stage=1))
    _assert_matches(matcher.match(d0, d1), oracle.match(d0, d1, stage=1))


@pytest.mark.parametrize("mcfg", [
    dict(DistanceType="HAMMING", FilterMatches=0, GoodMatchesCount=20, UseRatioTest=1, RatioTestThreshold=0.5),
    dict(DistanceType="HAMMING", FilterMatches=0, GoodMatchesCount=20, UseRatioTest=0, RatioTestThreshold=0.5),
    dict(DistanceType="HAMMING", FilterMatches=1, GoodMatchesCount=500, UseRatioTest=1, RatioTestThreshold=0.9),
    dict(DistanceType="HAMMING", FilterMatches=1, GoodMatchesCount=5, UseRatioTest=0, RatioTestThreshold=1.0),
    dict(DistanceType="HAMMING", FilterMatches=1, GoodMatchesCount=100000, UseRatioTest=0, RatioTestThreshold=1.0),
])
def test_match_configs(gpu_ctx, oracle, kitti_pair, tum_pair, mcfg):
    import slam_cin0051_b200 as s
    m = s.FeatureMatcher(mcfg, gpu_ctx)
    for pair in (kitti_pair, tum_pair):
        k0, d0 = oracle.detect_and_compute(pair[0])
        k1, d1 = oracle.detect_and_compute(pair[1])
        for with_kp in (True, False):
            want = oracle.match(d0, d1, k0 if with_kp else None, k1 if with_kp else None, cfg=mcfg, stage=1)
            got = m.match(d0, d1, k0 if with_kp else None, k1 if with_kp else None)
            _assert_matches(got, want)


def test_match_random_descriptors_and_edge_cases(gpu_ctx, matcher, oracle):
    import slam_cin0051_b200 as s
    rng = np.random.default_rng(0)
    m = s.FeatureMatcher(dict(DistanceType="HAMMING", FilterMatches=0, GoodMatchesCount=1, UseRatioTest=1,
                              RatioTestThreshold=0.95), gpu_ctx)
    for n1, n2, w in [(1, 1, 32), (3, 1, 32), (257, 130, 32), (100, 1000, 8), (64, 64, 5), (300, 300, 64)]:
        d1 = rng.integers(0, 256, (n1, w), dtype=np.uint8)
        d2 = rng.integers(0, 256, (n2, w), dtype=np.uint8)
        d2[: min(n1, n2) // 2] = d1[: min(n1, n2) // 2]  # exact duplicates -> distance-0 ties
        cfg = dict(FilterMatches=0, GoodMatchesCount=1, UseRatioTest=1, RatioTestThreshold=0.95)
        _assert_matches(m.match(d1, d2), oracle.match(d1, d2, cfg=cfg, stage=1))
    with pytest.raises(ValueError, match="Empty descriptors provided."):
        matcher.match(np.zeros((0, 32), np.uint8), np.zeros((4, 32), np.uint8))
    with pytest.raises(RuntimeError, match="Descriptor dimensions must match."):
        matcher.match(np.zeros((4, 32), np.uint8), np.zeros((4, 16), np.uint8))


def test_knn2_against_bruteforce(matcher):
    rng = np.random.default_rng(1)
    d1 = rng.integers(0, 256, (333, 32), dtype=np.uint8)
    d2 = rng.integers(0, 256, (777, 32), dtype=np.uint8)
    d2[5] = d2[700]
    d1[0] = d2[5]
    got = matcher.knn2(d1, d2)
    pop = np.array([bin(i).count("1") for i in range(256)], np.int32)
    dist = pop[d1[:, None, :] ^ d2[None, :, :]].sum(-1)
    order = np.lexsort((np.broadcast_to(np.arange(777), dist.shape), dist), axis=1)
    assert np.array_equal(got["trainIdx0"], order[:, 0])
    assert np.array_equal(got["trainIdx1"], order[:, 1])
    assert np.array_equal(got["distance0"], np.take_along_axis(dist, order[:, :1], 1)[:, 0])
    assert np.array_equal(got["distance1"], np.take_along_axis(dist, order[:, 1:2], 1)[:, 0])


def test_sequence_path_equals_single_calls(gpu_ctx, detector, matcher, oracle):
    import slam_cin0051_b200 as s
    from slam_cin0051_b200.synth import make_sequence
    frames = make_sequence(376, 1241, 6, pitch_px=14, seed=0)
    seq = s.FrameSequence(376, 1241, 6, context=gpu_ctx)
    seq.upload(frames)
    seq.extract(detector)
    seq.match_consecutive(matcher, with_keypoints=True)
    counts = seq.counts()
    assert (counts[:, 3] == 0).all()
    prev = None
    for f in range(6):
        wk, wd = oracle.detect_and_compute(frames[f])
        gk, gd = seq.frame(f)
        assert_kps_equal(gk, wk)
        assert np.array_equal(gd, wd)
        assert counts[f, 0] == len(wk)
        if prev is not None:
            _assert_matches(seq.matches(f - 1), oracle.match(prev[1], wd, prev[0], wk, stage=1))
        prev = (wk, wd)


def test_constructor_errors(gpu_ctx, tmp_path):
    import slam_cin0051_b200 as s
    with pytest.raises(RuntimeError, match="Could not open feature detector file"):
        s.FeatureDetector(tmp_path / "missing.yml", gpu_ctx)
    base = dict(IntensityThreshold=20, ContiguousPixelsThreshold=12, NonMaxSuppression=1, SuppressionWindowSize=12,
                PatchSize=31, NumBRIEFPairs=256)
    for key, bad, msg in [("IntensityThreshold", 300, "Intensity threshold"), ("ContiguousPixelsThreshold", 17, "Contiguous"),
                          ("NonMaxSuppression", 2, "Non-max"), ("SuppressionWindowSize", 0, "Suppression window"),
                          ("PatchSize", 30, "Patch size"), ("NumBRIEFPairs", 12, "BRIEF pairs")]:
        with pytest.raises(RuntimeError, match=msg):
            s.FeatureDetector({**base, key: bad}, gpu_ctx)
    with pytest.raises(RuntimeError, match="Invalid distance type"):
        s.FeatureMatcher(dict(DistanceType="COSINE"), gpu_ctx)
    l2 = s.FeatureMatcher(dict(DistanceType="L2", FilterMatches=0, UseRatioTest=0, RatioTestThreshold=0.5), gpu_ctx)
    with pytest.raises(RuntimeError, match="HAMMING"):
        l2.match(np.zeros((2, 32), np.uint8), np.zeros((2, 32), np.uint8))
