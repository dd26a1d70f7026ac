"""Randomised parity soak on a GPU box: random sizes, contents and parameters through the C ABI in both modes, compared
bit for bit with live cv2 (ORB mode, kNN) and with the C++ restatement of the reference (reference mode); RANSAC against
the oracle loop driven by the g++ build of the device solver and against cv2.findEssentialMat itself.
usage: python tools/soak_parity.py [seconds] [seed]
Prints one JSON line: cases run per family and every mismatch (seed + parameters, enough to reproduce).
tests/test_gpu_soak.py runs a bounded, fixed-seed slice of the same families inside `pytest -m gpu`.

Two-view tolerance: device == oracle loop (driven by the g++ build of the device's own solver) exactly; against cv2 as
stated in tests/ransac_compare.py (per problem: Jaccard >= 0.8, the mask is the exact Sampson test of the device's own E, and E
under a condition-aware bound wherever the masks are identical)."""
import ctypes as C
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import cv2  # noqa: E402

import slam_cin0051_b200 as S  # noqa: E402
from oracle import essential_oracle as eo  # noqa: E402
from oracle import ref_oracle  # noqa: E402
from slam_cin0051_b200.synth import make_sequence  # noqa: E402
from tools_golden import knn2 as cv_knn2  # noqa: E402
from tools_golden import orb_canonical  # noqa: E402

ctx = None
hx = None
kitti = None


def setup(context=None):
    global ctx, hx, kitti
    if ctx is not None:
        return
    ctx = context or S.Context(0)
    ref_oracle.build()
    so = os.path.join(ROOT, "tests", "native", "libhost_exact.so")
    if not os.path.exists(so):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-o", so, os.path.join(ROOT, "tests", "native", "host_exact.cpp")])
    hx = C.CDLL(so)
    kitti = cv2.imread(os.path.join(ROOT, "test", "data", "images", "0000000000.png"), 0)


def image(rng, rows, cols):
    kind = rng.integers(0, 4)
    if kind == 0:
        return make_sequence(rows, cols, 1, pitch_px=int(rng.integers(9, 30)), seed=int(rng.integers(1 << 30)))[0], "scene"
    if kind == 1:
        return rng.integers(0, 256, (rows, cols), dtype=np.uint8), "noise"
    if kind == 2 and rows <= kitti.shape[0] and cols <= kitti.shape[1]:
        y, x = rng.integers(0, kitti.shape[0] - rows + 1), rng.integers(0, kitti.shape[1] - cols + 1)
        return np.ascontiguousarray(kitti[y:y + rows, x:x + cols]), "kitti-crop"
    g = cv2.GaussianBlur(rng.integers(0, 256, (rows, cols), dtype=np.uint8), (0, 0), float(rng.uniform(0.8, 3.0)))
    return cv2.normalize(g, None, 0, 255, cv2.NORM_MINMAX).astype(np.uint8), "smooth-noise"


def orb_case(rng):
    rows, cols = int(rng.integers(70, 700)), int(rng.integers(70, 1300))
    p = dict(NumLevels=int(rng.integers(1, 9)), ScaleFactor=float(rng.choice([1.1, 1.2, 1.25, 1.3, 1.5, 2.0])),
             MaxFeatures=int(rng.choice([50, 300, 1000, 2000, 5000])), FastThreshold=int(rng.choice([5, 10, 20, 40])))
    img, kind = image(rng, rows, cols)
    det = S.FeatureDetector(dict(IntensityThreshold=20, ContiguousPixelsThreshold=9, NonMaxSuppression=1, SuppressionWindowSize=3,
                                 PatchSize=31, NumBRIEFPairs=256, **p), ctx)
    k, d = det.detect_and_compute(img)
    got = {"x": k["x"], "y": k["y"], "size": k["size"], "angle": k["angle"], "response": k["response"],
           "octave": det.last_octaves(len(k)), "desc": d}
    want = orb_canonical(img, nfeatures=p["MaxFeatures"], nlevels=p["NumLevels"], scale=p["ScaleFactor"], fast=p["FastThreshold"])
    ok = len(got["x"]) == len(want["x"])
    if ok:
        for f in ("x", "y", "size", "angle", "response"):
            ok &= np.asarray(got[f], np.float32).tobytes() == np.asarray(want[f], np.float32).tobytes()
        ok &= np.array_equal(got["octave"], want["octave"]) and np.array_equal(got["desc"].reshape(len(k), -1) if len(k) else np.zeros((0, 32), np.uint8),
                                                                                want["desc"].reshape(len(k), -1) if len(k) else np.zeros((0, 32), np.uint8))
    return bool(ok), dict(rows=rows, cols=cols, kind=kind, n=int(len(k)), **p)


def ref_case(rng):
    rows, cols = int(rng.integers(40, 300)), int(rng.integers(40, 500))  # the CPU restatement has the O(n^2) greedy NMS
    cfg = dict(IntensityThreshold=int(rng.choice([10, 20, 35])), ContiguousPixelsThreshold=int(rng.choice([0, 5, 9, 12, 16])),
               NonMaxSuppression=int(rng.integers(0, 2)), SuppressionWindowSize=int(rng.choice([3, 7, 12, 20])),
               PatchSize=int(rng.choice([9, 15, 31, 41])), NumBRIEFPairs=int(rng.choice([8, 64, 256, 512])))
    img, kind = image(rng, rows, cols)
    img2, _ = image(rng, rows, cols)
    det = S.FeatureDetector(cfg, ctx)
    gk, gd = det.detect_and_compute(img)
    wk, wd = ref_oracle.detect_and_compute(img, cfg)
    # no keypoints: the reference returns DescriptorMatrix(0, 0) (feature_detector.cpp:22-25), the oracle wrapper (0, 32)
    ok = gk.tobytes() == wk.tobytes() and (np.array_equal(gd, wd) if len(wk) else gd.size == 0)
    if ok and len(wk) > 0:
        wk2, wd2 = ref_oracle.detect_and_compute(img2, cfg)
        if len(wk2) > 0:
            mcfg = dict(DistanceType="HAMMING", FilterMatches=int(rng.integers(0, 2)), GoodMatchesCount=int(rng.choice([5, 20, 500])),
                        UseRatioTest=int(rng.integers(0, 2)), RatioTestThreshold=float(rng.choice([0.5, 0.75, 0.9])))
            m = S.FeatureMatcher(mcfg, ctx)
            with_kp = bool(rng.integers(0, 2))
            got = m.match(wd, wd2, wk if with_kp else None, wk2 if with_kp else None)
            q, t, d = ref_oracle.match(wd, wd2, wk if with_kp else None, wk2 if with_kp else None, cfg=mcfg, stage=1)
            ok = len(got) == len(q) and np.array_equal(got["queryIdx"], q) and np.array_equal(got["trainIdx"], t) and \
                np.array_equal(got["distance"].view(np.uint32), np.asarray(d, np.float32).view(np.uint32))
            cfg = {**cfg, **mcfg, "with_kp": with_kp}
    return bool(ok), dict(rows=rows, cols=cols, kind=kind, n=int(len(wk)), **cfg)


def knn_case(rng):
    nq, nt = int(rng.integers(1, 3000)), int(rng.integers(2, 3000))
    dq = rng.integers(0, 256, (nq, 32), dtype=np.uint8)
    dt = rng.integers(0, 256, (nt, 32), dtype=np.uint8)
    if rng.integers(0, 2):  # plant near-duplicates and exact ties
        k = min(nq, nt) // 2
        dt[:k] = dq[:k]
        dt[k // 2:k, 0] ^= 1
    mat = S.FeatureMatcher(dict(DistanceType="HAMMING", FilterMatches=0, GoodMatchesCount=1, UseRatioTest=1, RatioTestThreshold=0.75), ctx)
    got = mat.knn2(dq, dt)
    idx, dist = cv_knn2(dq, dt)
    ok = np.array_equal(got["trainIdx0"], idx[:, 0]) and np.array_equal(got["trainIdx1"], idx[:, 1]) and \
        np.array_equal(got["distance0"], dist[:, 0].astype(np.float32)) and np.array_equal(got["distance1"], dist[:, 1].astype(np.float32))
    return bool(ok), dict(nq=nq, nt=nt)


def solver(a, b):
    a = np.ascontiguousarray(a, np.float64).reshape(1, 5, 2)
    b = np.ascontiguousarray(b, np.float64).reshape(1, 5, 2)
    models = np.zeros((1, 10, 9))
    counts = np.zeros(1, np.int32)
    hx.hx_five_point(a.ctypes.data, b.ctypes.data, 1, models.ctypes.data, counts.ctypes.data)
    return [models[0, k].reshape(3, 3).copy() for k in range(counts[0])]


def ransac_case(rng):
    n = int(rng.integers(8, 400))
    inl = float(rng.uniform(0.3, 0.95))
    K4 = (float(rng.uniform(300, 900)),) * 2 + (320.0, 240.0)
    K = np.array([[K4[0], 0, K4[2]], [0, K4[1], K4[3]], [0, 0, 1.0]])
    X = np.c_[rng.uniform(-3, 3, n), rng.uniform(-2, 2, n), rng.uniform(4, 12, n)]
    rv = rng.normal(0, 0.08, 3)
    R, _ = cv2.Rodrigues(rv)
    t = rng.normal(0, 0.4, 3)
    p1 = (K @ X.T).T
    p2 = (K @ (X @ R.T + t).T).T
    p1, p2 = p1[:, :2] / p1[:, 2:], p2[:, :2] / p2[:, 2:]
    p2 = p2 + rng.normal(0, 0.4, p2.shape)
    bad = rng.random(n) > inl
    p2[bad] = rng.uniform(0, 640, (int(bad.sum()), 2))
    p1, p2 = p1.astype(np.float32), p2.astype(np.float32)
    max_iters = int(rng.choice([20, 56, 100, 1000]))
    E, mask, good = S.find_essential(p1, p2, K4, max_iters=max_iters, context=ctx)
    wE, wmask, wgood = eo.find_essential(p1, p2, K4, max_iters=max_iters, solver=solver)
    ok = good == wgood and np.array_equal(mask, wmask)  # device == the oracle loop driven by the g++ build of its solver: exact
    info = dict(n=n, inlier_ratio=round(inl, 2), max_iters=max_iters, good=int(good), want=int(wgood))
    if ok:
        from ransac_compare import compare
        ok, cinfo = compare(p1, p2, K4, E, mask, max_iters=max_iters)  # ... and cv2 itself, per-problem tolerance
        info.update(cinfo)
    return bool(ok), info


def sequence_case(rng):
    """Batched device-resident path and the pipelined host-buffer path against the single-image calls (all on the GPU)."""
    import torch
    from slam_cin0051_b200.sequence import FrameSequence
    orb = bool(rng.integers(0, 2))
    rows, cols, n = int(rng.integers(80, 420)), int(rng.integers(100, 700)), int(rng.integers(2, 10))
    frames = np.stack([image(rng, rows, cols)[0] if rng.integers(0, 4) else np.full((rows, cols), 90, np.uint8) for _ in range(n)])
    if orb:
        det = S.FeatureDetector(dict(IntensityThreshold=20, ContiguousPixelsThreshold=9, NonMaxSuppression=1, SuppressionWindowSize=3,
                                     PatchSize=31, NumBRIEFPairs=256, NumLevels=int(rng.integers(1, 7)), ScaleFactor=1.2,
                                     MaxFeatures=int(rng.choice([200, 1000])), FastThreshold=20), ctx)
        mat = S.FeatureMatcher(os.path.join(ROOT, "test", "data", "feature_matcher_orb.yml"), ctx)
    else:
        det = S.FeatureDetector(os.path.join(ROOT, "test", "data", "feature_detector.yml"), ctx)
        mat = S.FeatureMatcher(os.path.join(ROOT, "test", "data", "feature_matcher.yml"), ctx)
    cap, raw = 8192, rows * cols
    singles = [det.detect_and_compute(f) for f in frames]
    if max(len(k) for k, _ in singles) > cap:
        return True, dict(skipped="more keypoints than the sequence capacity")
    chunk = int(rng.choice([1, 2, 3, 64]))
    seq = FrameSequence(rows, cols, n, desc_bytes=32, max_raw_corners=raw, max_keypoints=cap, context=ctx)
    h_frames = torch.empty((n, rows, cols), dtype=torch.uint8, pin_memory=True)
    h_frames.numpy()[:] = frames
    h_kps = torch.zeros((n, cap, 5), dtype=torch.float32, pin_memory=True)
    h_desc = torch.zeros((n, cap, 32), dtype=torch.uint8, pin_memory=True)
    h_m = torch.zeros((n, cap, 3), dtype=torch.int32, pin_memory=True)
    h_c = torch.zeros((n, 4), dtype=torch.int32, pin_memory=True)
    seq.process_ptrs(det, mat, h_frames.data_ptr(), n, chunk=chunk, with_keypoints=not orb, kps_ptr=h_kps.data_ptr(),
                     desc_ptr=h_desc.data_ptr(), matches_ptr=h_m.data_ptr(), counts_ptr=h_c.data_ptr())
    ctx.synchronize()
    c = h_c.numpy()
    if (c[:, 3] != 0).any():
        return True, dict(skipped="candidate-list overflow reported by the sequence (status bits)")
    ok = True
    for f, (k, d) in enumerate(singles):
        ok &= int(c[f, 0]) == len(k) and h_kps.numpy()[f, :len(k)].tobytes() == k.tobytes()
        ok &= len(k) == 0 or np.array_equal(h_desc.numpy()[f, :len(k)], d)
        if f + 1 < n:
            k2, d2 = singles[f + 1]
            if len(k) and len(k2):
                m = mat.match(d, d2, None if orb else k, None if orb else k2)
                ok &= int(c[f, 1]) == len(m) and h_m.numpy()[f, :len(m)].tobytes() == m.tobytes()
            else:
                ok &= int(c[f, 1]) == 0
    return bool(ok), dict(orb=orb, rows=rows, cols=cols, n=n, chunk=chunk)


def prep_case(rng):
    """cv::cvtColor(BGR2GRAY) against cv2 and Camera::undistortImage against the C++ restatement."""
    rows, cols = int(rng.integers(8, 500)), int(rng.integers(8, 900))
    bgr = rng.integers(0, 256, (rows, cols, 3), dtype=np.uint8)
    ok = np.array_equal(S.bgr_to_gray(bgr, ctx), cv2.cvtColor(bgr, cv2.COLOR_BGR2GRAY))
    gray = rng.integers(0, 256, (rows, cols), dtype=np.uint8)
    K4 = (float(rng.uniform(200, 1200)), float(rng.uniform(200, 1200)), cols / 2 + float(rng.normal(0, 5)), rows / 2 + float(rng.normal(0, 5)))
    D4 = tuple(float(x) for x in rng.normal(0, [0.2, 0.05, 0.002, 0.002]))
    out = np.zeros((rows, cols), np.uint8)
    outf = np.zeros((rows, cols), np.float64)
    k = np.array(K4, np.float64)
    dd = np.array(D4, np.float64)
    ctx.check(ctx.lib.slamcu_undistort(ctx.handle, gray.ctypes.data, rows, cols, cols, k.ctypes.data, dd.ctypes.data, out.ctypes.data,
                                       outf.ctypes.data))
    want = ref_oracle.undistort(gray, K4, D4)  # the reference's double image in [0, 1]
    ok &= np.array_equal(outf, want) and np.array_equal(out.astype(np.float64) / 255.0, want)
    return bool(ok), dict(rows=rows, cols=cols, K4=K4, D4=D4)


families = {"orb_vs_cv2": orb_case, "reference_vs_cpp_restatement": ref_case, "knn2_vs_cv2": knn_case, "ransac_vs_oracle_loop_and_cv2": ransac_case,
            "sequence_vs_single_calls": sequence_case, "prep_vs_cv2_and_restatement": prep_case}


def run(budget=120.0, seed0=0, max_cases=None, context=None):
    """Round-robin over the families until `budget` seconds or `max_cases` cases; returns the JSON-able record."""
    setup(context)
    runs = {k: 0 for k in families}
    fails = []
    t_end = time.time() + budget
    i = 0
    while time.time() < t_end and (max_cases is None or i < max_cases):
        name = list(families)[i % len(families)]
        seed = seed0 * 1000003 + i
        try:
            ok, info = families[name](np.random.default_rng(seed))
        except Exception as e:  # a crash is a finding too
            ok, info = False, {"exception": repr(e)[:200]}
        runs[name] += 1
        if not ok:
            fails.append({"family": name, "seed": seed, **info})
        i += 1
    return {"tool": "soak_parity", "seconds": budget, "seed0": seed0, "cases": runs, "mismatches": len(fails), "details": fails[:20]}


if __name__ == "__main__":
    print(json.dumps(run(float(sys.argv[1]) if len(sys.argv) > 1 else 120.0, int(sys.argv[2]) if len(sys.argv) > 2 else 0)))
