"""GPU parity of the image-preparation rows (P1, P2 of SURVEY.md section 8a): cv::cvtColor(BGR2GRAY) as called at
src/preprocessing/preprocessor.cpp:136 and Camera::undistortImage (include/slam/common/common.hpp:127-173).
Bit-exact: gray bytes against cv2 itself; the undistorted image against the C++ restatement (both the u8 gather the
detector consumes and the reference's value / 255.0 double image)."""
import os

import numpy as np
import pytest

from conftest import DATA, load_gray

pytestmark = pytest.mark.gpu


def test_bgr_to_gray_equals_cv2(gpu_ctx):
    cv2 = pytest.importorskip("cv2")
    import slam_cin0051_b200 as s
    rng = np.random.default_rng(0)
    for shape in [(480, 640), (37, 53), (1, 1), (512, 1392)]:
        bgr = rng.integers(0, 256, shape + (3,), dtype=np.uint8)
        got = s.bgr_to_gray(bgr, gpu_ctx)
        assert np.array_equal(got, cv2.cvtColor(bgr, cv2.COLOR_BGR2GRAY)), shape
        b, g, r = (bgr[..., k].astype(np.int64) for k in range(3))
        assert np.array_equal(got, ((3735 * b + 19235 * g + 9798 * r + 16384) >> 15).astype(np.uint8))
    colour = cv2.imread(os.path.join(DATA, "test_images", "0.png"), cv2.IMREAD_COLOR)  # what Preprocessor::yield reads
    assert np.array_equal(s.bgr_to_gray(colour, gpu_ctx), cv2.cvtColor(colour, cv2.COLOR_BGR2GRAY))
    with pytest.raises(RuntimeError):
        s.bgr_to_gray(np.zeros((4, 4), np.uint8), gpu_ctx)


def test_undistort_equals_restatement(gpu_ctx, oracle):
    import slam_cin0051_b200 as s
    cam = s.Camera(os.path.join(DATA, "camera.yml"), 0, gpu_ctx)
    assert cam.image_size == (1392, 512) and cam.K[0, 0] == 984.2439 and len(cam.D) == 5
    img = load_gray("images/0000000000.png")
    K4 = [cam.fx, cam.fy, cam.cx, cam.cy]
    D4 = [cam.k1, cam.k2, cam.p1, cam.p2]  # k3 is loaded but unused by the reference (common.hpp:113,151-154)
    want, mp = oracle.undistort(img, K4, D4, want_map=True)
    got = cam.undistort_image(img)
    assert got.dtype == np.float64 and np.array_equal(got, want)
    u8 = cam.undistort_image_u8(img)
    inside = mp >= 0
    assert np.array_equal(u8[inside], img.reshape(-1)[mp[inside]]) and (u8[~inside] == 0).all()
    # errors mirrored from common.hpp:130-135
    with pytest.raises(RuntimeError, match="Input image is empty."):
        cam.undistort_image(np.zeros((0, 0), np.uint8))
    with pytest.raises(RuntimeError, match="does not match camera image size"):
        cam.undistort_image(img[:100])


def test_undistort_other_cameras(gpu_ctx, oracle, tmp_path):
    import slam_cin0051_b200 as s
    rng = np.random.default_rng(3)
    img = rng.integers(0, 256, (240, 320), dtype=np.uint8)
    for k1, k2, p1, p2 in [(0.0, 0.0, 0.0, 0.0), (0.25, -0.1, 0.01, -0.02), (-0.6, 0.4, 0.0, 0.0)]:
        yml = tmp_path / "cam.yml"
        yml.write_text("%YAML:1.0\n---\nImageSize: [320, 240]\n"
                       "K0: !!opencv-matrix\n   rows: 3\n   cols: 3\n   dt: d\n   data: [ 250., 0., 160.5, 0., 245., 119.5, 0., 0., 1. ]\n"
                       f"D0: !!opencv-matrix\n   rows: 4\n   cols: 1\n   dt: d\n   data: [ {k1}, {k2}, {p1}, {p2} ]\n")
        cam = s.Camera(yml, 0, gpu_ctx)
        want = oracle.undistort(img, [250.0, 245.0, 160.5, 119.5], [k1, k2, p1, p2])
        assert np.array_equal(cam.undistort_image(img), want), (k1, k2, p1, p2)
        if k1 > 0:  # pincushion: the forward map leaves the image near the border -> zeros (common.hpp:160-170)
            assert (want == 0).mean() > 0.02
    with pytest.raises(RuntimeError, match="Could not find keys"):
        s.Camera(yml, 3, gpu_ctx)


def test_sequence_prepare_batches_gray_and_undistort(gpu_ctx, oracle):
    """slamcu_sequence_prepare == per-frame cvtColor + undistortImage gather, and extraction runs on the prepared frames."""
    cv2 = pytest.importorskip("cv2")
    import slam_cin0051_b200 as s
    cam = s.Camera(os.path.join(DATA, "camera.yml"), 0, gpu_ctx)
    rng = np.random.default_rng(5)
    base = load_gray("images/0000000000.png")
    bgr = np.stack([np.stack([np.roll(base, k, 1), base, np.roll(base, -k, 0)], -1) for k in (1, 2, 3)])  # 3 colour frames
    bgr = (bgr.astype(np.int32) + rng.integers(-3, 4, bgr.shape)).clip(0, 255).astype(np.uint8)
    seq = s.FrameSequence(512, 1392, 3, desc_bytes=32, max_keypoints=4096, max_raw_corners=65536, context=gpu_ctx)
    K4, D4 = [cam.fx, cam.fy, cam.cx, cam.cy], [cam.k1, cam.k2, cam.p1, cam.p2]
    seq.prepare(bgr, cam)
    det = s.FeatureDetector(os.path.join(DATA, "feature_detector.yml"), gpu_ctx)
    seq.extract(det)
    for f in range(3):
        gray = cv2.cvtColor(bgr[f], cv2.COLOR_BGR2GRAY)
        _, mp = oracle.undistort(gray, K4, D4, want_map=True)
        want = np.where(mp >= 0, gray.reshape(-1)[np.maximum(mp, 0)], 0).astype(np.uint8)
        assert np.array_equal(seq.image(f), want), f
        wk, wd = oracle.detect_and_compute(want)
        gk, gd = seq.frame(f)
        assert gk.tobytes() == wk.tobytes() and np.array_equal(gd, wd)
    seq.prepare(bgr[:, :, :, 1].copy())  # gray, no camera: plain upload semantics
    assert np.array_equal(seq.image(1), bgr[1, :, :, 1])
