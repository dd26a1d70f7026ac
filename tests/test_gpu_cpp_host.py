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
    pat = np.load(os.path.join(ROOT, "slam_cin0051_b200", "orb_bit_pattern_31.npy")).astype("<i4")
    pat.tofile(tmp_path / "pattern.i32")
    cfg = open(os.path.join(DATA, "feature_detector_orb.yml")).read() + f'\nOrbPatternFile: "{tmp_path / "pattern.i32"}"\n'
    (tmp_path / "det.yml").write_text(cfg)
    out = subprocess.run([exe, str(tmp_path / "0.pgm"), str(tmp_path / "1.pgm"), str(tmp_path / "det.yml"),
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
