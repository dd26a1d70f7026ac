"""Pins the CPU oracle (oracle/ref_frontend.cpp) to the known answers for the reference's own fixtures.

The reference's tests assert nothing numeric ("parity unpinned" by the reference itself), so the pins are
the survey-time known answers recorded in SURVEY.md Appendix D for the reference's fixture images and
default YAML configuration: raw corner counts, kept keypoint counts, border (zero-descriptor) keypoints,
first keypoint, FNV-1a-64 digests of all descriptor bytes, the BRIEF pattern, and the match counts.
"""
import numpy as np

from conftest import load_gray


def test_brief_pattern_pin(oracle):
    pat = oracle.brief_pattern(31, 256)
    assert len(pat) == 46
    assert pat[:8].tolist() == [[1, -4, 7, 2], [-9, 3, 13, 4], [14, 0, 0, 14], [11, 11, -1, 1], [5, 14, 13, 14],
                                [-10, 6, -3, -7], [-8, -14, 4, 12], [-3, -9, -9, -6]]
    assert oracle.fnv1a64(bytes((pat.reshape(-1) + 64).astype(np.uint8))) == "a64782560e64c890"


def test_blur_weights_are_normalised(oracle):
    w = oracle.blur_weights()
    assert w.shape == (25,) and abs(w.sum() - 1.0) < 1e-15
    assert len(np.unique(np.round(w, 15))) == 6


PINS = {
    "images/0000000000.png": dict(shape=(512, 1392), raw=11329, kept=1145, border=81, first=(672.0, 187.0, 3114.0, -32.8727),
                                  digest="f03506a4c32e12dc"),
    "images/0000000001.png": dict(shape=(512, 1392), raw=11844, kept=1153, border=68),
    "test_images/0.png": dict(shape=(480, 640), raw=3692, kept=372, border=8, first=(517.0, 289.0, 2499.0, 12.7098),
                              digest="4fd496356eea6792"),
    "test_images/1.png": dict(shape=(480, 640), raw=1703, kept=219, border=21),
}


def test_fixture_known_answers(oracle):
    for rel, pin in PINS.items():
        img = load_gray(rel)
        assert img.shape == pin["shape"]
        assert len(oracle.fast_scan(img)) == pin["raw"]
        kps, desc = oracle.detect_and_compute(img)
        assert len(kps) == pin["kept"]
        r = 15
        border = ((kps["x"] - r < 0) | (kps["x"] + r >= img.shape[1]) | (kps["y"] - r < 0) | (kps["y"] + r >= img.shape[0]))
        assert int(border.sum()) == pin["border"]
        assert not desc[border].any() and (kps["angle"][border] == 0).all()
        assert not desc[:, 6:].any()  # only bits 0..45 can ever be set with the 46-pair pattern
        if "first" in pin:
            x, y, resp, ang = pin["first"]
            assert (kps["x"][0], kps["y"][0], kps["response"][0]) == (x, y, resp)
            assert abs(float(kps["angle"][0]) - ang) < 1e-4
            assert oracle.fnv1a64(desc.tobytes()) == pin["digest"]


def test_match_known_answers(oracle):
    k0, d0 = oracle.detect_and_compute(load_gray("images/0000000000.png"))
    k1, d1 = oracle.detect_and_compute(load_gray("images/0000000001.png"))
    q, t, d, pen = oracle.match(d0, d1, k0, k1, stage=0, return_penalised=True)
    assert len(q) == 90 and pen == 448456 and len(k0) * len(k1) == 1320185
    q2, t2, d2 = oracle.match(d0, d1, k0, k1, stage=1)
    assert len(q2) == 20 and (np.diff(d2) >= 0).all()
    t0, e0 = oracle.detect_and_compute(load_gray("test_images/0.png"))
    t1, e1 = oracle.detect_and_compute(load_gray("test_images/1.png"))
    assert len(oracle.match(e0, e1, stage=0)[0]) == 9  # >= 8, so PoseEstimator::estimate runs on this pair


def test_sort_is_unstable_like_libstdcxx(oracle):
    # the std::sort permutation differs from a stable sort on tie-heavy keys: the property the GPU must replicate
    rng = np.random.default_rng(0)
    resp = rng.integers(0, 50, 5000).astype(np.float32)
    perm = oracle.sort_perm_desc(resp)
    assert (np.diff(resp[perm]) <= 0).all()
    stable = np.argsort(-resp, kind="stable")
    assert not np.array_equal(perm, stable)


def test_undistort_matches_closed_form(oracle):
    img = load_gray("images/0000000000.png")
    K4 = [984.2439, 980.8141, 690.0, 233.1966]
    D4 = [-0.3728755, 0.2037299, 0.002219027, 0.001383707]
    out, mp = oracle.undistort(img, K4, D4, want_map=True)
    assert out.shape == img.shape and out.min() >= 0.0 and out.max() <= 1.0
    i, j = 233, 690  # principal point maps to itself
    assert mp[i, j] == i * img.shape[1] + j
    inside = mp >= 0
    assert np.array_equal(out[inside], img.reshape(-1)[mp[inside]] / 255.0)
    assert (out[~inside] == 0).all()
