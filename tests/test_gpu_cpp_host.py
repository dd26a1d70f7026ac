"""The C++ host side above the C ABI (include/slam/cuda/frontend.hpp adapters with the reference's class / method names,
tools/cli/slam_bench, test/frontend/test_frontend_cuda) against the oracle: the C++ program prints FNV-1a-64 digests of
its keypoints, descriptors and matches; they must equal the digests of the oracle's results on the same fixtures."""
import json
import os
import subprocess

import numpy as np
import pytest

from conftest import DATA, ROOT, load_gray

pytestmark = pytest.mark.gpu
BUILD = os.path.join(ROOT, "slam_cin0051_b200", "build")


def fnv(b: bytes) -> str:
    h = 1469598103934665603
    for x in b:
        h = ((h ^ x) * 1099511628211) & 0xFFFFFFFFFFFFFFFF
    return f"{h:016x}"


def write_pgm(path, img):
    with open(path, "wb") as f:
        f.write(b"P5\n%d %d\n255\n" % (img.shape[1], img.shape[0]))
        f.write(np.ascontiguousarray(img).tobytes())


def test_cpp_adapters_reference_mode(tmp_path, oracle):
    exe = os.path.join(BUILD, "test_frontend_cuda")
    assert os.path.exists(exe), "run __graft_entry__.build() first"
    a, b = load_gray("test_images/0.png"), load_gray("test_images/1.png")
    write_pgm(tmp_path / "0.pgm", a)
    write_pgm(tmp_path / "1.pgm", b)
    out = subprocess.run([exe, str(tmp_path / "0.pgm"), str(tmp_path / "1.pgm"), os.path.join(DATA, "feature_detector.yml"),
                          os.path.join(DATA, "feature_matcher.yml"), "525", "525", "319.5", "239.5"],
                         capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr
    lines = out.stdout.strip().splitlines()
    k0, d0 = oracle.detect_and_compute(a)
    k1, d1 = oracle.detect_and_compute(b)
    assert lines[0].split() == ["kp0", str(len(k0)), fnv(k0.tobytes()), "desc0", fnv(d0.tobytes())]
    assert lines[1].split() == ["kp1", str(len(k1)), fnv(k1.tobytes()), "desc1", fnv(d1.tobytes())]

    def mbytes(m):
        q, t, d = m
        rec = np.zeros(len(q), np.dtype([("q", "<i4"), ("t", "<i4"), ("d", "<f4")]))
        rec["q"], rec["t"], rec["d"] = q, t, d
        return rec.tobytes()
    mk = oracle.match(d0, d1, k0, k1, stage=1)
    mn = oracle.match(d0, d1, stage=1)
    assert lines[2].split() == ["match_kp", str(len(mk[0])), fnv(mbytes(mk)), "match_nokp", str(len(mn[0])), fnv(mbytes(mn))]
    assert lines[3] == "empty: Empty descriptors provided."
    assert lines[4].startswith("essential valid ")
    # 9 matches on this pair: the essential-matrix path runs on >= 8 correspondences
    assert len(mn[0]) == 9


def test_cpp_adapters_orb_mode(tmp_path):
    from test_orb_oracle import gold
    exe = os.path.join(BUILD, "test_frontend_cuda")
    a, b = load_gray("images/0000000000.png"), load_gray("images/0000000001.png")
    write_pgm(tmp_path / "0.pgm", a)
    write_pgm(tmp_path / "1.pgm", b)
    out = subprocess.run([exe, str(tmp_path / "0.pgm"), str(tmp_path / "1.pgm"), os.path.join(DATA, "feature_detector_orb.yml"),
                          os.path.join(DATA, "feature_matcher_orb.yml")], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr
    lines = out.stdout.strip().splitlines()
    for line, name in ((lines[0], "kitti0"), (lines[1], "kitti1")):
        g = gold(f"orb_{name}.npz")
        rec = np.zeros(len(g["x"]), np.dtype([(f, "<f4") for f in ("x", "y", "size", "angle", "response")]))
        for f in rec.dtype.names:
            rec[f] = g[f]
        tok = line.split()
        assert tok[1] == str(len(rec)) and tok[2] == fnv(rec.tobytes()) and tok[4] == fnv(g["desc"].tobytes())


def test_slam_bench_cli():
    exe = os.path.join(BUILD, "slam_bench")
    assert os.path.exists(exe)
    out = subprocess.run([exe, "-c", os.path.join(DATA, "feature_detector.yml"), "-m", os.path.join(DATA, "feature_matcher.yml"),
                          "-f", "32", "-s", "2", "-w", "3", "-k", "-C", "16"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr
    r = json.loads(out.stdout.strip().splitlines()[-1])
    assert r["frames_per_s_resident"] > 0 and r["frames_per_s_e2e"] > 0 and r["overflowed_frames"] == 0
    assert r["keypoints_per_frame"] > 500 and r["gpu_launches"] > 0


def test_cpp_camera_adapter(tmp_path, oracle):
    """slam::cuda::Camera (calibration YAML with !!opencv-matrix nodes, undistortImage) and bgrToGray through the C++
    adapters: digests equal the C++ restatement of Camera::undistortImage and cv2's BGR2GRAY; the reference's error
    messages for a size mismatch and a missing camera index."""
    cv2 = pytest.importorskip("cv2")
    exe = os.path.join(BUILD, "test_frontend_cuda")
    img = load_gray("images/0000000000.png")  # 1392 x 512, the size camera.yml declares
    write_pgm(tmp_path / "k.pgm", img)
    out = subprocess.run([exe, "--camera", os.path.join(DATA, "camera.yml"), "0", str(tmp_path / "k.pgm")],
                         capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr
    lines = out.stdout.strip().splitlines()
    from slam_cin0051_b200.config import read_yaml
    cfg = read_yaml(os.path.join(DATA, "camera.yml"))
    K = np.asarray(cfg["K0"], np.float64).reshape(3, 3)
    D = np.asarray(cfg["D0"], np.float64).reshape(-1)
    head = lines[0].split()
    assert head[:2] == ["camera", "1392x512"]
    assert np.array_equal(np.array(head[3:12], np.float64), K.ravel()) and np.array_equal(np.array(head[13:], np.float64), D)
    K4, D4 = (K[0, 0], K[1, 1], K[0, 2], K[1, 2]), tuple(D[:4])
    want = oracle.undistort(img, K4, D4)
    want8 = np.rint(want * 255.0).astype(np.uint8)
    assert lines[1].split() == ["undistort", "f64", fnv(want.tobytes()), "u8", fnv(want8.tobytes())]
    bgr = np.stack([img, 255 - img, img >> 1], -1)
    assert lines[2].split() == ["gray", fnv(cv2.cvtColor(bgr, cv2.COLOR_BGR2GRAY).tobytes())]
    assert lines[3] == "mismatch: Input image size does not match camera image size."
    assert lines[4] == "missing: Could not find keys K7 or D7 in file."


def test_cpp_pose_estimator_adapter(tmp_path):
    """slam::cuda::PoseEstimator with the reference's own signature -- estimate(pairs1, pairs2, matches, R, t) and
    triangulatePoints -- equals the Python mirror on the same ORB matches; R, t stay untouched below 8 matches."""
    import slam_cin0051_b200 as s
    import slam_cin0051_b200.pose as P
    exe = os.path.join(BUILD, "test_frontend_cuda")
    a, b = load_gray("images/0000000000.png"), load_gray("images/0000000001.png")
    write_pgm(tmp_path / "0.pgm", a)
    write_pgm(tmp_path / "1.pgm", b)
    K4 = (984.2439, 980.8141, 690.0, 233.1966)
    out = subprocess.run([exe, str(tmp_path / "0.pgm"), str(tmp_path / "1.pgm"), os.path.join(DATA, "feature_detector_orb.yml"),
                          os.path.join(DATA, "feature_matcher_orb.yml"), *[repr(v) for v in K4]], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr
    lines = {l.split()[0]: l.split()[1:] for l in out.stdout.strip().splitlines()}
    ctx = s.Context.default()
    det = s.FeatureDetector(os.path.join(DATA, "feature_detector_orb.yml"), ctx)
    mat = s.FeatureMatcher(os.path.join(DATA, "feature_matcher_orb.yml"), ctx)
    (k0, d0), (k1, d1) = det.detect_and_compute(a), det.detect_and_compute(b)
    m = mat.match(d0, d1)
    p1 = np.stack([k0["x"][m["queryIdx"]], k0["y"][m["queryIdx"]]], 1)
    p2 = np.stack([k1["x"][m["trainIdx"]], k1["y"][m["trainIdx"]]], 1)
    want = P.estimate_pose(p1, p2, K4, ctx)
    assert want is not None and len(m) >= 8
    tok = lines["estimator"]
    assert tok[:2] == ["touched", "1"] and tok[2] == "R" and tok[12] == "t"
    assert np.array_equal(np.array([float(v) for v in tok[3:12]]).reshape(3, 3), want["R"])
    assert np.array_equal(np.array([float(v) for v in tok[13:16]]), want["t"].ravel())
    assert lines["estimator_few"] == ["touched", "0"]
    K = np.array([[K4[0], 0, K4[2]], [0, K4[1], K4[3]], [0, 0, 1.0]])
    x3 = P.triangulate_points(K, want["R"], want["t"], p1, p2, ctx)
    tri = lines["triangulated"]
    assert int(tri[0]) == len(m)
    # (the adapter forms K [R | t] with its own operation order: the last bits of P2 differ from numpy's matmul)
    assert np.allclose(np.array([float(v) for v in tri[1:10]]).reshape(3, 3), x3[:3], rtol=1e-9, atol=1e-9)


def test_cpp_preprocessor_adapter(tmp_path):
    """slam::cuda::Preprocessor over a directory stream (file selection, lexical order, timestamps.txt, frameSkip): yield() equals
    cv2's BGR2GRAY followed by the restatement of Camera::undistortImage; the batched yieldInto() lands the same bytes in HBM."""
    cv2 = pytest.importorskip("cv2")
    from oracle import ref_oracle
    exe = os.path.join(BUILD, "test_frontend_cuda")
    rng = np.random.default_rng(3)
    d = tmp_path / "stream"
    d.mkdir()
    frames = []
    for i in range(5):  # .png / .jpg names carrying PPM payloads (the decoder is the adapter's template parameter)
        bgr = rng.integers(0, 256, (512, 1392, 3), dtype=np.uint8)
        frames.append(bgr)
        with open(d / f"{i:010d}.{'png' if i % 2 == 0 else 'jpg'}", "wb") as f:
            f.write(b"P6\n1392 512\n255\n" + bgr.tobytes())
    (d / "notes.txt").write_text("ignored")
    (d / "timestamps.txt").write_text("".join(f"2011-09-26 13:02:25.{964389445 + 103 * i:09d}\n" for i in range(5)))
    out = subprocess.run([exe, "--preprocess", str(d), os.path.join(DATA, "camera.yml")], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr
    lines = out.stdout.strip().splitlines()
    assert lines[0] == "frames 5"
    K4 = [9.842439e+02, 9.808141e+02, 6.900000e+02, 2.331966e+02]
    D4 = [-3.728755e-01, 2.037299e-01, 2.219027e-03, 1.383707e-03]
    # lexical order: .jpg and .png interleave by their numeric stem
    u8 = []
    for i, bgr in enumerate(frames):
        gray = cv2.cvtColor(bgr, cv2.COLOR_BGR2GRAY)
        und = ref_oracle.undistort(gray, K4, D4)
        u8.append(np.rint(und * 255.0).astype(np.uint8))
        tok = lines[1 + i].split()
        assert tok[:2] == ["yield", "512x1392"] and tok[2] == fnv(und.tobytes()), i
        assert int(tok[4]) == (964389445 + 103 * i) // 1_000_000  # the stamp truncated to milliseconds (preprocessor.cpp:114-119)
    assert lines[6] == "batched 3"  # frameSkip = 1: frames 0, 2, 4
    for slot, i in enumerate((0, 2, 4)):
        assert lines[7 + slot].split() == ["slot", str(slot), fnv(u8[i].tobytes())]
    assert lines[10].startswith("bad: Unsupported stream type: ")
