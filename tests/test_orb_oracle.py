"""CPU pins of the ORB-mode oracle (oracle/orb_oracle.py): every stage against vectors produced by OpenCV
itself -- the committed fixtures under tests/golden/ (tools/make_golden_orb.py) and, when cv2 is importable,
live cv2 calls.  Bit-exact: positions, sizes, octaves, Harris responses and angles (as bit patterns),
descriptor bytes, kNN indices and distances."""
import os

import numpy as np
import pytest

from conftest import ROOT, load_gray

GOLD = os.path.join(ROOT, "tests", "golden")


def gold(name):
    return np.load(os.path.join(GOLD, name))


def images():
    from slam_cin0051_b200.synth import make_sequence
    syn = make_sequence(376, 1241, 2, pitch_px=14, seed=0)
    return {"kitti0": load_gray("images/0000000000.png"), "kitti1": load_gray("images/0000000001.png"),
            "tum0": load_gray("test_images/0.png"), "synK0": syn[0], "synK1": syn[1]}


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def assert_orb_equal(got, want):
    assert len(got["x"]) == len(want["x"])
    for f in ("x", "y", "size", "angle", "response"):
        assert np.array_equal(bits(got[f]), bits(want[f])), f
    assert np.array_equal(np.asarray(got["octave"], np.int32), np.asarray(want["octave"], np.int32))
    if len(want["x"]):  # an empty result is DescriptorMatrix(0, 0) in the reference's convention
        assert np.array_equal(got["desc"], want["desc"])


def test_level_geometry_and_quotas():
    from oracle import orb_oracle as oo
    sc = oo.level_scales()
    assert oo.level_sizes(512, 1392, sc) == [(512, 1392), (427, 1160), (356, 967), (296, 806), (247, 671), (206, 559),
                                             (171, 466), (143, 388)]
    assert oo.level_quotas(2000) == [434, 362, 302, 251, 209, 175, 145, 122]
    assert sum(oo.level_quotas(1000)) == 1000 and sum(oo.level_quotas(10000)) == 10000


@pytest.mark.parametrize("name", ["kitti0", "tum0", "synK0"])
def test_oracle_equals_cv2_golden(name):
    from oracle import orb_oracle as oo
    r = oo.orb_detect_and_compute(images()[name])
    assert_orb_equal(r, gold(f"orb_{name}.npz"))
    assert len(r["x"]) == 2000


def test_oracle_small_parameters_golden():
    from oracle import orb_oracle as oo
    r = oo.orb_detect_and_compute(images()["tum0"], nfeatures=300, nlevels=4, scale_factor=1.5, fast_threshold=30)
    assert_orb_equal(r, gold("orb_tum0_n300_l4_s15_t30.npz"))


def test_knn2_equals_bfmatcher_golden():
    from oracle import orb_oracle as oo
    for a, b in (("kitti0", "kitti1"), ("synK0", "synK1")):
        g = gold(f"knn2_{a}_{b}.npz")
        idx, dist = oo.knn2(gold(f"orb_{a}.npz")["desc"], gold(f"orb_{b}.npz")["desc"])
        assert np.array_equal(idx, g["idx"]) and np.array_equal(dist.astype(np.float32), g["dist"])


def test_stages_against_live_cv2():
    cv2 = pytest.importorskip("cv2")
    from oracle import orb_oracle as oo
    img = images()["tum0"]
    lv1 = oo.resize_linear_exact(img, 400, 533)
    assert np.array_equal(lv1, cv2.resize(img, (533, 400), interpolation=cv2.INTER_LINEAR_EXACT))
    fast = cv2.FastFeatureDetector_create(threshold=20, nonmaxSuppression=True, type=cv2.FAST_FEATURE_DETECTOR_TYPE_9_16)
    kp = fast.detect(img)
    xs, ys, sc = oo.fast9_detect(img, 20)
    assert [(int(k.pt[0]), int(k.pt[1]), int(k.response)) for k in kp] == list(zip(xs.tolist(), ys.tolist(), sc.tolist()))
    rng = np.random.default_rng(0)
    y = rng.integers(-2 ** 22, 2 ** 22, 20000).astype(np.float32)
    x = rng.integers(-2 ** 22, 2 ** 22, 20000).astype(np.float32)
    want = np.array([cv2.fastAtan2(float(a), float(b)) for a, b in zip(y[:3000], x[:3000])], np.float32)
    assert np.array_equal(bits(oo.fast_atan2(y[:3000], x[:3000])), bits(want))
    # the golden fixtures are what the installed cv2 still produces
    from tools_golden import orb_canonical
    assert_orb_equal(orb_canonical(img), gold("orb_tum0.npz"))


def test_level_sizes_use_the_reciprocal_scale():
    """ORB_Impl sizes level l as cvRound(dim * (1.0f / scale)), not cvRound(dim / scale): 477 px at scale 1.2f is
    397.49997 as a quotient (-> 397) and exactly 397.5 as a product (-> 398, half to even).  Checked against cv2 itself:
    its octave-1 keypoints are FAST corners of the 398-wide level, not of the 397-wide one."""
    cv2 = pytest.importorskip("cv2")
    from oracle import orb_oracle as oo
    scales = oo.level_scales(2, 1.2)
    assert oo.level_sizes(306, 477, scales)[1] == (255, 398)
    assert oo.level_sizes(200, 117, scales)[1][1] == 98 and oo.level_sizes(376, 1241, scales)[1] == (313, 1034)
    rng = np.random.default_rng(0)
    img = cv2.GaussianBlur(rng.integers(0, 256, (200, 477), dtype=np.uint8), (0, 0), 1.2)
    orb = cv2.ORB_create(nfeatures=3000, scaleFactor=1.2, nlevels=2, edgeThreshold=31, firstLevel=0, WTA_K=2,
                         scoreType=cv2.ORB_HARRIS_SCORE, patchSize=31, fastThreshold=10)
    pts = {(int(round(k.pt[0] / 1.2000000476837158)), int(round(k.pt[1] / 1.2000000476837158))) for k in orb.detect(img, None) if k.octave == 1}
    assert len(pts) > 100
    lv = oo.build_pyramid(img, 2, 1.2)[0][1]
    assert lv.shape == (167, 398)
    corners = {(int(p.pt[0]), int(p.pt[1])) for p in cv2.FastFeatureDetector_create(10, True).detect(lv)}
    assert pts <= corners
