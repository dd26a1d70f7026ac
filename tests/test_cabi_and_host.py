"""CPU checks of the boundary: the C-ABI library loads and exports every symbol include/slam/cuda/slamcu.h
declares, the product fails loudly without a GPU (no CPU fallback), the YAML reader handles the
reference's configuration files, and the host-table helpers agree with the oracle."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from conftest import DATA, ROOT


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "slam", "cuda", "slamcu.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(slamcu_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    import slam_cin0051_b200 as S
    from slam_cin0051_b200 import _lib
    lib = S.load()
    declared = _declared_symbols()
    assert len(declared) >= 35
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in slamcu.h but not exported"
        assert name in _lib.SYMBOLS, f"{name} has no ctypes signature"
    assert lib.slamcu_abi_version() == 1
    assert lib.slamcu_status_string(2) == b"Empty descriptors provided."


def test_no_cpu_fallback_without_gpu():
    import slam_cin0051_b200 as S
    lib = S.load()
    n = C.c_int(0)
    lib.slamcu_device_count(C.byref(n))
    if n.value > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(S.SlamcuError, match="no CUDA device"):
        S.Context(0)
    with pytest.raises(S.SlamcuError):
        S.FeatureDetector(os.path.join(DATA, "feature_detector.yml"))
    h = C.c_void_p()
    assert lib.slamcu_create(0, C.byref(h)) != 0 and not h.value


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "slam_cin0051_b200")
    for dirpath, _, files in os.walk(pkg):
        if os.path.basename(dirpath) == "build":
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                text = open(os.path.join(dirpath, f)).read()
                assert "ref_oracle" not in text and "libslam_oracle" not in text and "oracle/" not in text, f


def test_host_tables_match_oracle(oracle):
    import slam_cin0051_b200 as S
    lib = S.load()
    for patch, pairs in [(31, 256), (15, 64), (41, 512), (9, 8)]:
        pat = np.zeros((pairs, 4), np.int32)
        n = C.c_int(0)
        assert lib.slamcu_default_brief_pattern(patch, pairs, pat.ctypes.data, pairs, C.byref(n)) == 0
        assert np.array_equal(pat[: n.value], oracle.brief_pattern(patch, pairs))
    w = np.zeros(25, np.float64)
    assert lib.slamcu_default_blur_weights(w.ctypes.data) == 0
    assert np.array_equal(w, oracle.blur_weights())


def test_yaml_reader_on_reference_configs(tmp_path):
    from slam_cin0051_b200.config import get_float, get_int, get_str, read_yaml
    det = read_yaml(os.path.join(DATA, "feature_detector.yml"))
    assert [get_int(det, k) for k in ("IntensityThreshold", "ContiguousPixelsThreshold", "NonMaxSuppression",
                                      "SuppressionWindowSize", "PatchSize", "NumBRIEFPairs")] == [20, 12, 1, 12, 31, 256]
    mat = read_yaml(os.path.join(DATA, "feature_matcher.yml"))
    assert get_str(mat, "DistanceType") == "HAMMING" and get_int(mat, "GoodMatchesCount") == 20
    assert get_float(mat, "RatioTestThreshold") == 0.5
    cam = read_yaml(os.path.join(DATA, "camera.yml"))
    assert cam["ImageSize"] == [1392, 512]
    assert cam["K0"].shape == (3, 3) and cam["K0"][0, 0] == 984.2439 and cam["D0"].shape == (5, 1)
    assert get_int(det, "Missing") == 0  # cv::FileNode >> int on a missing node leaves 0
    with pytest.raises(RuntimeError):
        read_yaml(tmp_path / "nope.yml")
    try:
        import cv2
    except ImportError:
        return
    fs = cv2.FileStorage(os.path.join(DATA, "camera.yml"), cv2.FILE_STORAGE_READ)
    assert np.array_equal(fs.getNode("K0").mat(), cam["K0"]) and np.array_equal(fs.getNode("D0").mat(), cam["D0"])


def test_synthetic_sequence_is_deterministic_and_translating():
    from slam_cin0051_b200.synth import make_sequence
    a = make_sequence(120, 200, 4, pitch_px=14, seed=3)
    b = make_sequence(120, 200, 4, pitch_px=14, seed=3)
    assert a.dtype == np.uint8 and a.shape == (4, 120, 200) and np.array_equal(a, b)
    # frame f is the canvas shifted by (2f, f): true correspondences between consecutive frames
    assert np.array_equal(a[0][1:, 2:], a[1][:-1, :-2])


def test_cmake_and_python_builds_list_the_same_sources():
    from slam_cin0051_b200 import build
    text = open(os.path.join(ROOT, "CMakeLists.txt")).read()
    listed = set(re.findall(r"/(\w+\.cu)\b", text))
    assert listed == set(build.SOURCES)
    assert set(f for f in os.listdir(build.CSRC) if f.endswith(".cu")) == set(build.SOURCES)
    assert "100a" in text and "-fmad=false" in text


def test_both_build_recipes_pass_the_exactness_flags():
    """Every bit-exactness guarantee rests on -fmad=false and on the absence of fast-math, in BOTH build recipes."""
    from slam_cin0051_b200 import build as b
    assert "-fmad=false" in b.NVCC_FLAGS and not any("fast_math" in f or "fast-math" in f for f in b.NVCC_FLAGS)
    cm = open(os.path.join(ROOT, "CMakeLists.txt")).read()
    assert "-fmad=false" in cm and "fast_math" not in cm and "fast-math" not in cm


def test_builtin_orb_pattern_is_the_extracted_cv2_table():
    """csrc/orb_pattern.inc (compiled into the library; used when slamcu_detector_config.orb_pattern is NULL) holds exactly the
    256 x 4 table tools/extract_orb_pattern.py recovered from cv2 (slam_cin0051_b200/orb_bit_pattern_31.npy)."""
    txt = open(os.path.join(ROOT, "slam_cin0051_b200", "csrc", "orb_pattern.inc")).read()
    vals = [int(t) for line in txt.splitlines() if not line.lstrip().startswith("//") for t in line.replace(",", " ").split()]
    want = np.load(os.path.join(ROOT, "slam_cin0051_b200", "orb_bit_pattern_31.npy")).astype(int).reshape(-1)
    assert len(vals) == 1024 and vals == want.tolist()
