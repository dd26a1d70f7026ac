"""ctypes front door to oracle/_ref/libslam_ref.so -- TEST INFRASTRUCTURE ONLY.

libslam_ref.so is the reference's OWN frontend: the unmodified /root/reference/src/frontend/feature_detector.cpp and
feature_matcher.cpp (+ include/slam/common/common.hpp's Camera), compiled by `make -C oracle ref` against the header
stand-ins in oracle/shim/ (Eigen / OpenCV-core / spdlog are not in this image).  It exists to pin the C++ restatement
(oracle/ref_frontend.cpp): tests/test_ref_build.py compares the two.  /root/reference exists only in the build container;
the GPU box gets the prebuilt .so with the snapshot (oracle/_ref/ is git-ignored, not gpurun-ignored).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import tempfile

import numpy as np

from .ref_oracle import DEFAULT_DET, DEFAULT_MAT, KP_DTYPE, _img, _p

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_ref", "libslam_ref.so")
REFERENCE = os.environ.get("SLAM_REFERENCE_DIR", "/root/reference")


def available() -> bool:
    return os.path.exists(_SO) or os.path.isdir(os.path.join(REFERENCE, "src", "frontend"))


def build(force: bool = False) -> str:
    """Compiles the reference sources where they lie; a no-op when /root/reference is absent and the .so is there."""
    if os.path.isdir(os.path.join(REFERENCE, "src", "frontend")):
        subprocess.check_call(["make", "-C", _HERE, "-s", "ref", f"REF={REFERENCE}"] + (["-B"] if force else []))
    if not os.path.exists(_SO):
        raise RuntimeError("oracle/_ref/libslam_ref.so is missing and /root/reference is not here to build it from")
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(build())
        L.ref_last_error.restype = C.c_char_p
        _lib = L
    return _lib


def last_error() -> str:
    return lib().ref_last_error().decode()


class Yml:
    """A temporary OpenCV-style YAML file (the reference's constructors take a path)."""

    def __init__(self, d: dict):
        self.f = tempfile.NamedTemporaryFile("w", suffix=".yml", delete=False)
        self.f.write("%YAML:1.0\n---\n")
        for k, v in d.items():
            self.f.write(f'{k}: "{v}"\n' if isinstance(v, str) else f"{k}: {v}\n")
        self.f.close()
        self.path = self.f.name.encode()

    def __del__(self):
        try:
            os.unlink(self.f.name)
        except OSError:
            pass


def det_yml(cfg=None) -> Yml:
    return Yml({**DEFAULT_DET, **(cfg or {})})


def mat_yml(cfg=None) -> Yml:
    return Yml({"DistanceType": "HAMMING", **DEFAULT_MAT, **(cfg or {})})


def _check(n):
    if n < 0:
        raise RuntimeError(f"reference call failed ({n}): {last_error()}")
    return n


def brief_pattern(cfg=None):
    y = det_yml(cfg)
    out = np.zeros((4096, 4), np.int32)
    n = _check(lib().ref_brief_pattern(y.path, _p(out, C.c_int), len(out)))
    return out[:n].copy()


def fast_scan(img, cfg=None):
    img = _img(img)
    y = det_yml(cfg)
    out = np.zeros(img.size, KP_DTYPE)
    n = _check(lib().ref_fast_scan(y.path, _p(img, C.c_uint8), img.shape[0], img.shape[1], C.c_void_p(out.ctypes.data), len(out)))
    return out[:n].copy()


def gaussian_blur(img, ksize=5, sigma=1.0):
    img = _img(img)
    out = np.zeros_like(img)
    _check(lib().ref_gaussian_blur(_p(img, C.c_uint8), img.shape[0], img.shape[1], ksize, C.c_double(sigma), _p(out, C.c_uint8)))
    return out


def detect(img, cfg=None):
    img = _img(img)
    y = det_yml(cfg)
    out = np.zeros(img.size, KP_DTYPE)
    n = _check(lib().ref_detect(y.path, _p(img, C.c_uint8), img.shape[0], img.shape[1], C.c_void_p(out.ctypes.data), len(out)))
    return out[:n].copy()


def compute(img, kps, cfg=None):
    img = _img(img)
    y = det_yml(cfg)
    kps = np.ascontiguousarray(kps, KP_DTYPE).copy()
    nb = int({**DEFAULT_DET, **(cfg or {})}["NumBRIEFPairs"]) // 8
    desc = np.zeros((len(kps), nb), np.uint8)
    _check(lib().ref_compute(y.path, _p(img, C.c_uint8), img.shape[0], img.shape[1], C.c_void_p(kps.ctypes.data), len(kps),
                             _p(desc, C.c_uint8)))
    return kps, desc


def detect_and_compute(img, cfg=None):
    img = _img(img)
    y = det_yml(cfg)
    nb = int({**DEFAULT_DET, **(cfg or {})}["NumBRIEFPairs"]) // 8
    kps = np.zeros(img.size, KP_DTYPE)
    desc = np.zeros((img.size, nb), np.uint8)
    n = _check(lib().ref_detect_and_compute(y.path, _p(img, C.c_uint8), img.shape[0], img.shape[1], C.c_void_p(kps.ctypes.data),
                                            _p(desc, C.c_uint8), len(kps)))
    return kps[:n].copy(), desc[:n].copy()


def match(d1, d2, kp1=None, kp2=None, cfg=None):
    """slam::FeatureMatcher(cfg).match(d1, d2, out, kp1, kp2): returns (queryIdx, trainIdx, distance)."""
    d1 = np.ascontiguousarray(d1, np.uint8)
    d2 = np.ascontiguousarray(d2, np.uint8)
    y = mat_yml(cfg)
    k1 = np.ascontiguousarray(kp1, KP_DTYPE) if kp1 is not None else np.zeros(0, KP_DTYPE)
    k2 = np.ascontiguousarray(kp2, KP_DTYPE) if kp2 is not None else np.zeros(0, KP_DTYPE)
    n1 = len(d1)
    q, t, d = np.zeros(max(n1, 1), np.int32), np.zeros(max(n1, 1), np.int32), np.zeros(max(n1, 1), np.float32)
    w1 = d1.shape[1] if d1.ndim == 2 else 0
    w2 = d2.shape[1] if d2.ndim == 2 else 0
    n = _check(lib().ref_match(y.path, _p(d1, C.c_uint8), n1, w1, _p(d2, C.c_uint8), len(d2), w2, C.c_void_p(k1.ctypes.data), len(k1),
                               C.c_void_p(k2.ctypes.data), len(k2), _p(q, C.c_int), _p(t, C.c_int), _p(d, C.c_float), len(q)))
    return q[:n].copy(), t[:n].copy(), d[:n].copy()


def hamming(a, b):
    a = np.ascontiguousarray(a, np.uint8)
    b = np.ascontiguousarray(b, np.uint8)
    return lib().ref_hamming(_p(a, C.c_uint8), _p(b, C.c_uint8), len(a))


def undistort(img, camera_yml: str, camera_index: int = 0):
    img = _img(img)
    out = np.zeros(img.shape, np.float64)
    _check(lib().ref_undistort(camera_yml.encode(), camera_index, _p(img, C.c_uint8), img.shape[0], img.shape[1], _p(out, C.c_double)))
    return out


def detector_error(cfg) -> str:
    """'' when the reference constructor accepts the configuration, else '<exception type>: <message>'."""
    y = Yml(cfg)
    return "" if lib().ref_detector_check(y.path) == 0 else last_error()


def matcher_error(cfg) -> str:
    y = Yml(cfg)
    return "" if lib().ref_matcher_check(y.path) == 0 else last_error()
