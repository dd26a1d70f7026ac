"""ctypes front door to oracle/libslam_oracle.so -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module.  The product package (slam_cin0051_b200) never does.

The .so is the C++ restatement of the reference frontend (oracle/ref_frontend.cpp); it is built by
`make -C oracle` (done by __graft_entry__.build()).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libslam_oracle.so")

KP_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"), ("response", "<f4")])

DEFAULT_DET = dict(IntensityThreshold=20, ContiguousPixelsThreshold=12, NonMaxSuppression=1,
                   SuppressionWindowSize=12, PatchSize=31, NumBRIEFPairs=256)
DEFAULT_MAT = dict(FilterMatches=1, GoodMatchesCount=20, UseRatioTest=1, RatioTestThreshold=0.5)


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "ref_frontend.cpp")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        u8p, i32p, f32p, f64p = (C.POINTER(C.c_uint8), C.POINTER(C.c_int), C.POINTER(C.c_float),
                                 C.POINTER(C.c_double))
        L.orc_brief_pattern.argtypes = [C.c_int, C.c_int, i32p, C.c_int]
        L.orc_blur_weights.argtypes = [C.c_int, C.c_double, f64p]
        L.orc_fast_scan.argtypes = [u8p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int]
        L.orc_fast_scan_scored.argtypes = L.orc_fast_scan.argtypes
        L.orc_detect.argtypes = [u8p, C.c_int, C.c_int, i32p, C.c_void_p, C.c_int]
        L.orc_gaussian_blur.argtypes = [u8p, C.c_int, C.c_int, C.c_int, C.c_double, u8p]
        L.orc_compute.argtypes = [u8p, C.c_int, C.c_int, i32p, C.c_void_p, C.c_int, u8p]
        L.orc_detect_and_compute.argtypes = [u8p, C.c_int, C.c_int, i32p, C.c_void_p, u8p, C.c_int]
        L.orc_match.argtypes = [u8p, C.c_int, u8p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int,
                                i32p, C.c_float, C.c_int, i32p, i32p, f32p, C.c_int, C.POINTER(C.c_longlong)]
        L.orc_sort_perm_desc.argtypes = [f32p, C.c_int, i32p]
        L.orc_topk_perm_asc.argtypes = [f32p, C.c_int, C.c_int, i32p]
        L.orc_atan2f.argtypes = [f32p, f32p, f32p, C.c_longlong]
        L.orc_sincosf.argtypes = [f32p, f32p, f32p, C.c_longlong]
        L.orc_undistort.argtypes = [u8p, C.c_int, C.c_int, f64p, f64p, f64p, i32p]
        L.orc_frontend_run.argtypes = [u8p, C.c_int, C.c_int, C.c_int, i32p, i32p, C.c_float, C.c_int, C.c_int, i32p]
        L.orc_frontend_run.restype = C.c_double
        _lib = L
    return _lib


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def _img(a):
    a = np.ascontiguousarray(a, dtype=np.uint8)
    assert a.ndim == 2
    return a


def det_cfg(cfg=None):
    d = dict(DEFAULT_DET)
    d.update(cfg or {})
    return np.array([d["IntensityThreshold"], d["ContiguousPixelsThreshold"], d["NonMaxSuppression"],
                     d["SuppressionWindowSize"], d["PatchSize"], d["NumBRIEFPairs"]], dtype=np.int32)


def mat_cfg(cfg=None):
    d = dict(DEFAULT_MAT)
    d.update(cfg or {})
    return (np.array([d["FilterMatches"], d["GoodMatchesCount"], d["UseRatioTest"]], dtype=np.int32),
            float(d["RatioTestThreshold"]))


def brief_pattern(patch=31, pairs=256):
    out = np.zeros((pairs, 4), dtype=np.int32)
    n = lib().orc_brief_pattern(patch, pairs, _p(out, C.c_int), pairs)
    return out[:n].copy()


def blur_weights(ksize=5, sigma=1.0):
    out = np.zeros(ksize * ksize, dtype=np.float64)
    lib().orc_blur_weights(ksize, sigma, _p(out, C.c_double))
    return out


def fast_scan(img, thr=20, arc=12, scored=False):
    img = _img(img)
    cap = img.size
    out = np.zeros(cap, dtype=KP_DTYPE)
    fn = lib().orc_fast_scan_scored if scored else lib().orc_fast_scan
    n = fn(_p(img, C.c_uint8), img.shape[0], img.shape[1], thr, arc, out.ctypes.data, cap)
    return out[:n].copy()


def detect(img, cfg=None):
    img = _img(img)
    cap = img.size
    out = np.zeros(cap, dtype=KP_DTYPE)
    c = det_cfg(cfg)
    n = lib().orc_detect(_p(img, C.c_uint8), img.shape[0], img.shape[1], _p(c, C.c_int), out.ctypes.data, cap)
    return out[:n].copy()


def gaussian_blur(img, ksize=5, sigma=1.0):
    img = _img(img)
    out = np.zeros_like(img)
    lib().orc_gaussian_blur(_p(img, C.c_uint8), img.shape[0], img.shape[1], ksize, sigma, _p(out, C.c_uint8))
    return out


def compute(img, kps, cfg=None):
    img = _img(img)
    c = det_cfg(cfg)
    kps = np.ascontiguousarray(kps, dtype=KP_DTYPE).copy()
    nb = int(c[5]) // 8
    desc = np.zeros((len(kps), nb), dtype=np.uint8)
    if len(kps):
        lib().orc_compute(_p(img, C.c_uint8), img.shape[0], img.shape[1], _p(c, C.c_int), kps.ctypes.data, len(kps),
                          _p(desc, C.c_uint8))
    return kps, desc


def detect_and_compute(img, cfg=None):
    img = _img(img)
    c = det_cfg(cfg)
    cap = img.size
    nb = int(c[5]) // 8
    kps = np.zeros(cap, dtype=KP_DTYPE)
    desc = np.zeros((cap, nb), dtype=np.uint8)
    n = lib().orc_detect_and_compute(_p(img, C.c_uint8), img.shape[0], img.shape[1], _p(c, C.c_int), kps.ctypes.data,
                                     _p(desc, C.c_uint8), cap)
    return kps[:n].copy(), desc[:n].copy()


def match(d1, d2, kp1=None, kp2=None, cfg=None, stage=1, return_penalised=False):
    """stage 0: after the ratio test, query order; stage 1: after filterAndSortMatches."""
    d1 = np.ascontiguousarray(d1, dtype=np.uint8)
    d2 = np.ascontiguousarray(d2, dtype=np.uint8)
    assert d1.shape[1] == d2.shape[1]
    c, ratio = mat_cfg(cfg)
    n1, n2 = len(d1), len(d2)
    k1 = np.ascontiguousarray(kp1, dtype=KP_DTYPE) if kp1 is not None else np.zeros(0, KP_DTYPE)
    k2 = np.ascontiguousarray(kp2, dtype=KP_DTYPE) if kp2 is not None else np.zeros(0, KP_DTYPE)
    q = np.zeros(n1, np.int32)
    t = np.zeros(n1, np.int32)
    d = np.zeros(n1, np.float32)
    pen = C.c_longlong(0)
    n = lib().orc_match(_p(d1, C.c_uint8), n1, _p(d2, C.c_uint8), n2, d1.shape[1], k1.ctypes.data, len(k1),
                        k2.ctypes.data, len(k2), _p(c, C.c_int), ratio, stage, _p(q, C.c_int), _p(t, C.c_int),
                        _p(d, C.c_float), n1, C.byref(pen))
    res = (q[:n].copy(), t[:n].copy(), d[:n].copy())
    return res + (pen.value,) if return_penalised else res


def sort_perm_desc(resp):
    resp = np.ascontiguousarray(resp, dtype=np.float32)
    perm = np.zeros(len(resp), np.int32)
    lib().orc_sort_perm_desc(_p(resp, C.c_float), len(resp), _p(perm, C.c_int))
    return perm


def topk_perm_asc(dist, k):
    dist = np.ascontiguousarray(dist, dtype=np.float32)
    perm = np.zeros(max(len(dist), 1), np.int32)
    m = lib().orc_topk_perm_asc(_p(dist, C.c_float), len(dist), k, _p(perm, C.c_int))
    return perm[:m].copy()


def atan2f(y, x):
    y = np.ascontiguousarray(y, np.float32)
    x = np.ascontiguousarray(x, np.float32)
    out = np.zeros_like(y)
    lib().orc_atan2f(_p(y, C.c_float), _p(x, C.c_float), _p(out, C.c_float), y.size)
    return out


def sincosf(a):
    a = np.ascontiguousarray(a, np.float32)
    s = np.zeros_like(a)
    c = np.zeros_like(a)
    lib().orc_sincosf(_p(a, C.c_float), _p(s, C.c_float), _p(c, C.c_float), a.size)
    return s, c


def undistort(img, K4, D4, want_map=False):
    img = _img(img)
    K4 = np.ascontiguousarray(K4, np.float64)
    D4 = np.ascontiguousarray(D4, np.float64)
    out = np.zeros(img.shape, np.float64)
    mp = np.zeros(img.shape, np.int32) if want_map else None
    lib().orc_undistort(_p(img, C.c_uint8), img.shape[0], img.shape[1], _p(K4, C.c_double), _p(D4, C.c_double),
                        _p(out, C.c_double), _p(mp, C.c_int) if want_map else None)
    return (out, mp) if want_map else out


def frontend_run(frames, det=None, mat=None, with_kp=True, threads=1):
    """Times detectAndCompute on every frame + match(f, f+1).  Returns (seconds, counts[n,3])."""
    frames = np.ascontiguousarray(frames, dtype=np.uint8)
    n, rows, cols = frames.shape
    dc = det_cfg(det)
    mc, ratio = mat_cfg(mat)
    counts = np.zeros((n, 3), np.int32)
    sec = lib().orc_frontend_run(_p(frames, C.c_uint8), n, rows, cols, _p(dc, C.c_int), _p(mc, C.c_int), ratio,
                                 int(with_kp), threads, _p(counts, C.c_int))
    return sec, counts


def fnv1a64(data: bytes) -> str:
    h = 1469598103934665603
    for b in data:
        h ^= b
        h = (h * 1099511628211) & 0xFFFFFFFFFFFFFFFF
    return f"{h:016x}"
