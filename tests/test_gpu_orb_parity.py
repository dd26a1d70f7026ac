"""GPU parity of the ORB-compatible mode through the C ABI.

Bit-exact against (i) the committed cv2 outputs under tests/golden/, (ii) the numpy oracle stage by stage
(pyramid bytes, FAST-9 candidates + scores, retainBest sets, Harris responses), (iii) live cv2 when importable.
Float fields are compared as bit patterns.  cv2's in-level keypoint order is implementation defined, so the
contract is the canonical (octave, y, x) order, which is the order the CUDA path emits.
"""
import os

import numpy as np
import pytest

from conftest import DATA
from test_orb_oracle import assert_orb_equal, bits, gold, images

pytestmark = pytest.mark.gpu

ORB_CFG = dict(IntensityThreshold=20, ContiguousPixelsThreshold=9, NonMaxSuppression=1, SuppressionWindowSize=3,
               PatchSize=31, NumBRIEFPairs=256, NumLevels=8, ScaleFactor=1.2, MaxFeatures=2000)


@pytest.fixture(scope="module")
def orb_det(gpu_ctx):
    import slam_cin0051_b200 as s
    return s.FeatureDetector(os.path.join(DATA, "feature_detector_orb.yml"), gpu_ctx)


@pytest.fixture(scope="module")
def orb_mat(gpu_ctx):
    import slam_cin0051_b200 as s
    return s.FeatureMatcher(os.path.join(DATA, "feature_matcher_orb.yml"), gpu_ctx)


def gpu_orb(det, img):
    k, d = det.detect_and_compute(img)
    return {"x": k["x"], "y": k["y"], "size": k["size"], "angle": k["angle"], "response": k["response"],
            "octave": det.last_octaves(len(k)), "desc": d}


@pytest.mark.parametrize("name", ["kitti0", "kitti1", "tum0", "synK0", "synK1"])
def test_detect_and_compute_equals_cv2_golden(orb_det, name):
    r = gpu_orb(orb_det, images()[name])
    assert_orb_equal(r, gold(f"orb_{name}.npz"))
    key = np.stack([r["octave"].astype(np.float64), r["y"], r["x"]], 1)
    assert (np.lexsort((key[:, 2], key[:, 1], key[:, 0])) == np.arange(len(key))).all(), "output must be in canonical order"


def test_stage_probes_equal_oracle(orb_det):
    from oracle import orb_oracle as oo
    img = images()["tum0"]
    orb_det.detect_and_compute(img)
    levels, scales = oo.build_pyramid(img)
    quotas = oo.level_quotas(2000)
    for l, lv in enumerate(levels):
        assert np.array_equal(orb_det.orb_level_image(l), lv), f"pyramid level {l}"
        H, W = lv.shape
        xs, ys, sc = oo.fast9_detect(lv, 20)
        inb = (xs >= 31) & (xs < W - 31) & (ys >= 31) & (ys < H - 31)
        xs, ys, sc = xs[inb], ys[inb], sc[inb]
        gx, gy, gv = orb_det.orb_stage(0, l)
        assert np.array_equal(gx, xs) and np.array_equal(gy, ys) and np.array_equal(gv.astype(np.int32), sc), f"FAST level {l}"
        keep = oo.retain_best(sc.astype(np.float32), 2 * quotas[l])
        hr = oo.harris_responses(oo.reflect101(lv, 32), 32, xs[keep], ys[keep])
        gx, gy, gv = orb_det.orb_stage(1, l)
        assert np.array_equal(gx, xs[keep]) and np.array_equal(gy, ys[keep]) and np.array_equal(bits(gv), bits(hr)), f"Harris level {l}"
        keep2 = oo.retain_best(hr, quotas[l])
        gx, gy, gv = orb_det.orb_stage(2, l)
        assert np.array_equal(gx, xs[keep][keep2]) and np.array_equal(gy, ys[keep][keep2]), f"retainBest level {l}"
        ext = oo.reflect101(lv, 32)
        assert np.array_equal(orb_det.orb_level_image(l, blurred=True), oo.blur7_level(ext, 32, H, W)), f"blur level {l}"


def test_other_orb_parameters(gpu_ctx):
    import slam_cin0051_b200 as s
    det = s.FeatureDetector({**ORB_CFG, "NumLevels": 4, "ScaleFactor": 1.5, "MaxFeatures": 300, "FastThreshold": 30}, gpu_ctx)
    assert_orb_equal(gpu_orb(det, images()["tum0"]), gold("orb_tum0_n300_l4_s15_t30.npz"))


def test_against_live_cv2_on_other_inputs(gpu_ctx, orb_det):
    pytest.importorskip("cv2")
    import slam_cin0051_b200 as s
    from tools_golden import orb_canonical
    from slam_cin0051_b200.synth import make_sequence
    rng = np.random.default_rng(5)
    cases = [make_sequence(480, 640, 1, pitch_px=17, seed=2)[0],            # TUM-shape synthetic
             make_sequence(376, 1241, 3, pitch_px=14, seed=9)[2],
             rng.integers(0, 256, (200, 333), dtype=np.uint8),              # white noise: every level saturates its quota
             np.full((240, 320), 77, np.uint8)]                             # flat: no corners at all
    for img in cases:
        assert_orb_equal(gpu_orb(orb_det, img), orb_canonical(img))
    # 477 columns: level 1 is 398 wide (cvRound(477 * (1.0f / 1.2f)) = cvRound(397.5)), not 397 (found by tools/soak_parity.py)
    det5 = s.FeatureDetector({**ORB_CFG, "NumLevels": 5, "MaxFeatures": 1000}, gpu_ctx)
    odd = make_sequence(306, 477, 1, pitch_px=15, seed=31)[0]
    assert_orb_equal(gpu_orb(det5, odd), orb_canonical(odd, nfeatures=1000, nlevels=5))
    # more levels than the image can hold: the tiny ones have no keypoints (31-px border) and are not processed
    small = make_sequence(90, 140, 1, pitch_px=11, seed=32)[0]
    assert_orb_equal(gpu_orb(orb_det, small), orb_canonical(small))
    # white noise at a low threshold: a quarter of the pixels are FAST candidates
    det_lo = s.FeatureDetector({**ORB_CFG, "FastThreshold": 5, "MaxFeatures": 5000}, gpu_ctx)
    assert_orb_equal(gpu_orb(det_lo, cases[2]), orb_canonical(cases[2], nfeatures=5000, fast=5))
    det = s.FeatureDetector({**ORB_CFG, "MaxFeatures": 10000}, gpu_ctx)
    big = make_sequence(1080, 1920, 1, pitch_px=28, seed=4)[0]
    assert_orb_equal(gpu_orb(det, big), orb_canonical(big, nfeatures=10000))
    k, d = orb_det.detect_and_compute(cases[3])
    assert len(k) == 0 and d.shape == (0, 0)


@pytest.mark.parametrize("scale,levels", [(1.1, 8), (1.3, 6), (1.45, 5), (2.0, 4)])
def test_pyramid_scale_factors_against_live_cv2(gpu_ctx, scale, levels):
    """Both pyramid kernels (shared-memory tiles while the source window of a 128x64 tile fits, the per-pixel gather for
    steeper factors) over sizes that leave partial tiles on both axes."""
    pytest.importorskip("cv2")
    import slam_cin0051_b200 as s
    from tools_golden import orb_canonical
    from slam_cin0051_b200.synth import make_sequence
    det = s.FeatureDetector({**ORB_CFG, "NumLevels": levels, "ScaleFactor": scale, "MaxFeatures": 1500}, gpu_ctx)
    for rows, cols, seed in ((376, 1241, 21), (517, 903, 22), (130, 257, 23)):
        img = make_sequence(rows, cols, 1, pitch_px=13, seed=seed)[0]
        assert_orb_equal(gpu_orb(det, img), orb_canonical(img, nfeatures=1500, nlevels=levels, scale=scale))


def test_knn2_and_ratio_matches(orb_mat):
    for a, b in (("kitti0", "kitti1"), ("synK0", "synK1")):
        g = gold(f"knn2_{a}_{b}.npz")
        d1, d2 = gold(f"orb_{a}.npz")["desc"], gold(f"orb_{b}.npz")["desc"]
        got = orb_mat.knn2(d1, d2)
        assert np.array_equal(got["trainIdx0"], g["idx"][:, 0]) and np.array_equal(got["trainIdx1"], g["idx"][:, 1])
        assert np.array_equal(bits(got["distance0"]), bits(g["dist"][:, 0]))
        assert np.array_equal(bits(got["distance1"]), bits(g["dist"][:, 1]))
        # Lowe ratio 0.75 on the BFMatcher result, query order (feature_matcher_orb.yml)
        keep = ~(g["dist"][:, 0] >= np.float32(0.75) * g["dist"][:, 1])
        m = orb_mat.match(d1, d2)
        assert np.array_equal(m["queryIdx"], np.nonzero(keep)[0]) and np.array_equal(m["trainIdx"], g["idx"][keep, 0])
        assert np.array_equal(bits(m["distance"]), bits(g["dist"][keep, 0]))


def test_sequence_path_equals_golden(gpu_ctx, orb_det, orb_mat):
    import slam_cin0051_b200 as s
    from slam_cin0051_b200.synth import make_sequence
    frames = make_sequence(376, 1241, 4, pitch_px=14, seed=0)
    seq = s.FrameSequence(376, 1241, 4, desc_bytes=32, max_keypoints=2560, context=gpu_ctx)
    seq.upload(frames)
    seq.extract(orb_det)
    seq.match_consecutive(orb_mat, with_keypoints=False)
    counts = seq.counts()
    assert (counts[:, 3] == 0).all() and (counts[:, 0] == 2000).all()
    for f, name in ((0, "synK0"), (1, "synK1")):
        k, d = seq.frame(f)
        r = {"x": k["x"], "y": k["y"], "size": k["size"], "angle": k["angle"], "response": k["response"],
             "octave": seq.octaves(f, len(k)), "desc": d}
        assert_orb_equal(r, gold(f"orb_{name}.npz"))
    g = gold("knn2_synK0_synK1.npz")
    keep = ~(g["dist"][:, 0] >= np.float32(0.75) * g["dist"][:, 1])
    m = seq.matches(0)
    assert np.array_equal(m["queryIdx"], np.nonzero(keep)[0]) and np.array_equal(m["trainIdx"], g["idx"][keep, 0])
    # frames 2, 3 against the single-frame path
    for f in (2, 3):
        k, d = seq.frame(f)
        k1, d1 = orb_det.detect_and_compute(frames[f])
        assert k.tobytes() == k1.tobytes() and np.array_equal(d, d1)


def test_orb_mode_errors(gpu_ctx):
    import slam_cin0051_b200 as s
    with pytest.raises(RuntimeError, match="ORB mode requires"):
        s.FeatureDetector({**ORB_CFG, "NumBRIEFPairs": 128}, gpu_ctx)
    det = s.FeatureDetector(ORB_CFG, gpu_ctx)
    # levels below 8 px are not processed (nothing in them survives the 31-px border): no keypoints, like cv2
    k, d = det.detect_and_compute(np.zeros((20, 20), np.uint8))
    assert len(k) == 0
    with pytest.raises(s.SlamcuError):
        det.detect_and_compute(np.zeros((5, 40), np.uint8))  # the image itself is below the FAST support
