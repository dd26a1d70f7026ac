"""Pins the C++ restatement (oracle/ref_frontend.cpp) to THE REFERENCE'S OWN CODE.

oracle/_ref/libslam_ref.so is the unmodified /root/reference/src/frontend/feature_detector.cpp + feature_matcher.cpp
(+ Camera::undistortImage from include/slam/common/common.hpp) compiled against the header stand-ins in oracle/shim/
(`make -C oracle ref`).  Every `-m gpu` parity test of reference mode compares the CUDA path with the restatement; this
file closes the chain: restatement == reference, bit for bit, on the reference's fixtures, on the detector / matcher
configurations the GPU tests use, on noise, and on the constructors' error behaviour.
"""
import os

import numpy as np
import pytest

from conftest import DATA, load_gray

from oracle import ref_build as R

pytestmark = pytest.mark.skipif(not R.available(), reason="neither /root/reference nor a prebuilt oracle/_ref/libslam_ref.so")

FIXTURES = ["images/0000000000.png", "images/0000000001.png", "test_images/0.png", "test_images/1.png"]
DET_CONFIGS = [
    dict(IntensityThreshold=20, ContiguousPixelsThreshold=9, NonMaxSuppression=1, SuppressionWindowSize=7, PatchSize=31, NumBRIEFPairs=256),
    dict(IntensityThreshold=35, ContiguousPixelsThreshold=12, NonMaxSuppression=0, SuppressionWindowSize=12, PatchSize=31, NumBRIEFPairs=256),
    dict(IntensityThreshold=10, ContiguousPixelsThreshold=16, NonMaxSuppression=1, SuppressionWindowSize=20, PatchSize=15, NumBRIEFPairs=64),
    dict(IntensityThreshold=25, ContiguousPixelsThreshold=0, NonMaxSuppression=1, SuppressionWindowSize=12, PatchSize=41, NumBRIEFPairs=512),
    dict(IntensityThreshold=15, ContiguousPixelsThreshold=5, NonMaxSuppression=1, SuppressionWindowSize=3, PatchSize=9, NumBRIEFPairs=8),
]
MAT_CONFIGS = [
    dict(FilterMatches=1, GoodMatchesCount=20, UseRatioTest=1, RatioTestThreshold=0.5),
    dict(FilterMatches=0, GoodMatchesCount=20, UseRatioTest=1, RatioTestThreshold=0.5),
    dict(FilterMatches=0, GoodMatchesCount=20, UseRatioTest=0, RatioTestThreshold=0.5),
    dict(FilterMatches=1, GoodMatchesCount=500, UseRatioTest=1, RatioTestThreshold=0.9),
    dict(FilterMatches=1, GoodMatchesCount=5, UseRatioTest=0, RatioTestThreshold=1.0),
    dict(FilterMatches=1, GoodMatchesCount=100000, UseRatioTest=0, RatioTestThreshold=1.0),
]


def _same_kps(a, b):
    assert len(a) == len(b)
    assert a.tobytes() == b.tobytes()  # x, y, size, angle, response: all five floats bit for bit


def test_reference_builds_from_its_own_sources():
    so = R.build()
    assert os.path.exists(so)
    if os.path.isdir("/root/reference"):  # the recipe compiles the sources where they lie: nothing is copied into the repo
        mk = open(os.path.join(os.path.dirname(so), "..", "Makefile")).read()
        assert "$(REF)/src/frontend/feature_detector.cpp" in mk and "$(REF)/src/frontend/feature_matcher.cpp" in mk


def test_brief_pattern_and_known_answers(oracle):
    for cfg in [None] + DET_CONFIGS:
        c = {**R.DEFAULT_DET, **(cfg or {})}
        assert np.array_equal(R.brief_pattern(cfg), oracle.brief_pattern(c["PatchSize"], c["NumBRIEFPairs"]))
    pat = R.brief_pattern()
    assert len(pat) == 46 and oracle.fnv1a64(bytes((pat.reshape(-1) + 64).astype(np.uint8))) == "a64782560e64c890"
    img = load_gray(FIXTURES[0])
    assert len(R.fast_scan(img)) == 11329
    k, d = R.detect_and_compute(img)
    assert len(k) == 1145 and oracle.fnv1a64(d.tobytes()) == "f03506a4c32e12dc"


@pytest.mark.parametrize("rel", FIXTURES)
def test_fixtures_default_config(oracle, rel):
    img = load_gray(rel)
    _same_kps(R.fast_scan(img), oracle.fast_scan(img))
    assert np.array_equal(R.gaussian_blur(img), oracle.gaussian_blur(img))
    _same_kps(R.detect(img), oracle.detect(img))
    rk, rd = R.detect_and_compute(img)
    wk, wd = oracle.detect_and_compute(img)
    _same_kps(rk, wk)
    assert np.array_equal(rd, wd)


@pytest.mark.parametrize("cfg", DET_CONFIGS)
def test_detector_configs(oracle, cfg):
    img = load_gray(FIXTURES[3])
    rk, rd = R.detect_and_compute(img, cfg)
    wk, wd = oracle.detect_and_compute(img, cfg)
    _same_kps(rk, wk)
    assert np.array_equal(rd, wd)


def test_noise_flat_and_tiny_images(oracle):
    rng = np.random.default_rng(5)
    noise = rng.integers(0, 256, (97, 131), dtype=np.uint8)
    smooth = np.clip(rng.normal(128, 3, (90, 120)), 0, 255).astype(np.uint8)  # many blur sums near .5
    for img in (noise, smooth, np.full((40, 50), 77, np.uint8), noise[:7, :9], noise[:6, :40]):
        assert np.array_equal(R.gaussian_blur(img), oracle.gaussian_blur(img))
        for cfg in (None, DET_CONFIGS[0], DET_CONFIGS[4]):
            rk, rd = R.detect_and_compute(img, cfg)
            wk, wd = oracle.detect_and_compute(img, cfg)
            _same_kps(rk, wk)
            assert rd.shape == wd.shape and np.array_equal(rd, wd)


def test_compute_on_caller_keypoints(oracle):
    img = load_gray(FIXTURES[2])
    rng = np.random.default_rng(3)
    kps = np.zeros(400, R.KP_DTYPE)
    kps["x"] = rng.integers(0, img.shape[1], 400)
    kps["y"] = rng.integers(0, img.shape[0], 400)
    kps["size"] = 6.0
    rk, rd = R.compute(img, kps)
    wk, wd = oracle.compute(img, kps)
    _same_kps(rk, wk)
    assert np.array_equal(rd, wd)


@pytest.mark.parametrize("mcfg", MAT_CONFIGS)
def test_matcher_configs(oracle, mcfg):
    for a, b in ((FIXTURES[0], FIXTURES[1]), (FIXTURES[2], FIXTURES[3])):
        k0, d0 = oracle.detect_and_compute(load_gray(a))
        k1, d1 = oracle.detect_and_compute(load_gray(b))
        for with_kp in (True, False):
            kk = (k0, k1) if with_kp else (None, None)
            rq, rt, rd = R.match(d0, d1, *kk, cfg=mcfg)
            wq, wt, wd = oracle.match(d0, d1, *kk, cfg=mcfg, stage=1)
            assert np.array_equal(rq, wq) and np.array_equal(rt, wt) and rd.tobytes() == wd.tobytes()


def test_matcher_ties_ragged_widths_and_hamming(oracle):
    rng = np.random.default_rng(0)
    cfg = dict(FilterMatches=0, GoodMatchesCount=1, UseRatioTest=1, RatioTestThreshold=0.95)
    for n1, n2, w in [(1, 1, 32), (3, 1, 32), (257, 130, 32), (100, 1000, 8), (64, 64, 5), (300, 300, 64)]:
        d1 = rng.integers(0, 256, (n1, w), dtype=np.uint8)
        d2 = rng.integers(0, 256, (n2, w), dtype=np.uint8)
        d2[: min(n1, n2) // 2] = d1[: min(n1, n2) // 2]  # exact duplicates: distance-0 ties
        for c in (cfg, MAT_CONFIGS[0], MAT_CONFIGS[4]):
            r, w_ = R.match(d1, d2, cfg=c), oracle.match(d1, d2, cfg=c, stage=1)
            assert all(np.array_equal(x, y) for x, y in zip(r, w_))
    pop = np.array([bin(i).count("1") for i in range(256)])
    for _ in range(50):
        a, b = rng.integers(0, 256, (2, 32), dtype=np.uint8)
        assert R.hamming(a, b) == int(pop[a ^ b].sum())


def test_error_behaviour_matches_the_mirrors():
    base = dict(R.DEFAULT_DET)
    assert R.detector_error(base) == ""
    # the C ABI (api.cu) and the C++ adapters (frontend.hpp) carry the reference's messages verbatim
    from conftest import ROOT
    api = open(os.path.join(ROOT, "slam_cin0051_b200", "csrc", "api.cu")).read()
    hpp = open(os.path.join(ROOT, "include", "slam", "cuda", "frontend.hpp")).read()
    for key, bad in [("IntensityThreshold", 300), ("ContiguousPixelsThreshold", 17), ("NonMaxSuppression", 2), ("SuppressionWindowSize", 0),
                     ("PatchSize", 30), ("NumBRIEFPairs", 12)]:
        err = R.detector_error({**base, key: bad})
        assert err.startswith("runtime_error: ")
        msg = err[len("runtime_error: "):]
        assert f'"{msg}"' in api and f'"{msg}"' in hpp, msg
    mbase = {"DistanceType": "HAMMING", **R.DEFAULT_MAT}
    assert R.matcher_error(mbase) == ""
    assert R.matcher_error({**mbase, "DistanceType": "COSINE"}) == "runtime_error: Invalid distance type. Must be 'HAMMING' or 'L2'."
    assert "FilterMatches must be either 0" in R.matcher_error({**mbase, "FilterMatches": 2})
    assert "GoodMatchesCount must be positive" in R.matcher_error({**mbase, "GoodMatchesCount": 0})
    assert "RatioTestThreshold must be in the range" in R.matcher_error({**mbase, "RatioTestThreshold": 1.5})
    with pytest.raises(RuntimeError, match="invalid_argument: Empty descriptors provided."):
        R.match(np.zeros((0, 32), np.uint8), np.zeros((4, 32), np.uint8))
    with pytest.raises(RuntimeError, match="runtime_error: Descriptor dimensions must match."):
        R.match(np.zeros((4, 32), np.uint8), np.zeros((4, 16), np.uint8))
    with pytest.raises(RuntimeError, match="requires HAMMING"):
        R.match(np.zeros((4, 32), np.uint8), np.zeros((4, 32), np.uint8), cfg={"DistanceType": "L2"})


def test_undistort_image(oracle):
    cam = os.path.join(DATA, "camera.yml")
    img = load_gray(FIXTURES[0])  # 1392 x 512 = camera.yml's ImageSize
    K4 = [9.842439e+02, 9.808141e+02, 6.900000e+02, 2.331966e+02]
    D4 = [-3.728755e-01, 2.037299e-01, 2.219027e-03, 1.383707e-03]
    got = R.undistort(img, cam)
    want = oracle.undistort(img, K4, D4)
    assert got.tobytes() == want.tobytes()
    with pytest.raises(RuntimeError, match="Input image size does not match camera image size."):
        R.undistort(img[:100], cam)
