"""numpy restatement of OpenCV's ORB + BFMatcher(k=2) path -- TEST INFRASTRUCTURE ONLY.

"Mode B" of the frontend: what BASELINE.json's headline config names
(8-level pyramid, FAST-9 + Harris, nfeatures budget, rBRIEF-256, brute-force kNN).  The algorithm lives in
OpenCV (features2d/src/orb.cpp, fast.cpp, imgproc resize.cpp / filter.simd.hpp), an un-vendored dependency of
the reference (conanfile.txt:2 pins opencv/4.12.0; this image carries the cv2 4.13.0 wheel, no sources).
Every stage below restates OpenCV's published algorithm and is pinned against cv2 itself in
tests/test_orb_oracle.py (bit-exact keypoints, responses, angles, descriptors, kNN results).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this module.
"""
from __future__ import annotations

import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
F32 = np.float32

HARRIS_K = F32(0.04)
EDGE = 31
PATCH = 31
HALF_PATCH = 15


def bit_pattern_31():
    return np.load(os.path.join(_HERE, "..", "slam_cin0051_b200", "orb_bit_pattern_31.npy"))


# ---- B.1 / B.2: scales, level sizes, per-level quotas -------------------------------------------
def level_scales(nlevels=8, scale_factor=1.2):
    sf = np.float64(F32(scale_factor))
    return np.array([F32(np.power(sf, float(l))) for l in range(nlevels)], F32)


def level_sizes(rows, cols, scales):
    out = []
    for s in scales:
        inv = F32(1.0) / F32(s)  # ORB_Impl: inv_scale = 1.0f / scale; Size(cvRound(cols * inv_scale), cvRound(rows * inv_scale))
        w = int(np.rint(F32(cols) * inv))  # cvRound(float): round half to even.  NOT cols / scale: the two differ for
        h = int(np.rint(F32(rows) * inv))  # 81 (size, level) combinations below 2200 px (e.g. 477 -> 398, not 397)
        out.append((h, w))
    return out


def level_quotas(nfeatures, nlevels=8, scale_factor=1.2):
    factor = F32(1.0 / np.float64(F32(scale_factor)))
    nd = F32(nfeatures) * (F32(1) - factor) / (F32(1) - F32(np.power(np.float64(factor), float(nlevels))))
    q, total = [], 0
    for _ in range(nlevels - 1):
        v = int(np.rint(nd))
        q.append(v)
        total += v
        nd = F32(nd * factor)
    q.append(max(nfeatures - total, 0))
    return q


# ---- B.3: INTER_LINEAR_EXACT resize (fixed point 8.8 coefficients) -------------------------------
def _coeffs(dst, src):
    scale = 1.0 / (float(dst) / float(src))
    d = np.arange(dst, dtype=np.float64)
    f = scale * (d + 0.5) - 0.5
    i = np.floor(f).astype(np.int64)
    a = np.rint((f - i) * 256.0).astype(np.int64)
    i0 = i.copy()
    i1 = i + 1
    lo = i < 0
    hi = i >= src - 1
    i0[lo] = 0
    i1[lo] = 0
    a[lo] = 0
    i0[hi] = src - 1
    i1[hi] = src - 1
    a[hi] = 0
    return i0, i1, a


def resize_linear_exact(img, rows, cols):
    src = img.astype(np.int64)
    x0, x1, ax = _coeffs(cols, img.shape[1])
    y0, y1, ay = _coeffs(rows, img.shape[0])
    h = (256 - ax)[None, :] * src[:, x0] + ax[None, :] * src[:, x1]
    v = (256 - ay)[:, None] * h[y0, :] + ay[:, None] * h[y1, :]
    return ((v + 32768) >> 16).astype(np.uint8)


def build_pyramid(img, nlevels=8, scale_factor=1.2):
    scales = level_scales(nlevels, scale_factor)
    sizes = level_sizes(img.shape[0], img.shape[1], scales)
    levels = [np.ascontiguousarray(img)]
    for l in range(1, nlevels):
        levels.append(resize_linear_exact(levels[l - 1], sizes[l][0], sizes[l][1]))
    return levels, scales


def reflect101(img, border):
    return np.pad(img, border, mode="reflect")


# ---- B.4: FAST-9/16 with score and 3x3 NMS --------------------------------------------------------
_RING = [(0, 3), (1, 3), (2, 2), (3, 1), (3, 0), (3, -1), (2, -2), (1, -3), (0, -3), (-1, -3), (-2, -2), (-3, -1),
         (-3, 0), (-3, 1), (-2, 2), (-1, 3)]  # (dx, dy)


def fast9_scores(img, threshold):
    """Score map (0 = not a corner): cornerScore = max over the 16 arcs of 9 of min |diff|, minus 1."""
    H, W = img.shape
    I = img.astype(np.int32)
    c = I[3:H - 3, 3:W - 3]
    d = np.stack([c - I[3 + dy:H - 3 + dy, 3 + dx:W - 3 + dx] for dx, dy in _RING], 0)  # v - I_k
    d2 = np.concatenate([d, d[:8]], 0)
    dark = np.full(c.shape, -10**6, np.int32)
    bright = np.full(c.shape, -10**6, np.int32)
    for s in range(16):
        arc = d2[s:s + 9]
        dark = np.maximum(dark, arc.min(0))
        bright = np.maximum(bright, (-arc).min(0))
    m = np.maximum(dark, bright)
    score = np.zeros((H, W), np.int32)
    score[3:H - 3, 3:W - 3] = np.where(m > threshold, m - 1, 0)
    return score


def fast9_detect(img, threshold=20):
    """Returns (x, y, score) int arrays in raster order after 3x3 non-max suppression."""
    s = fast9_scores(img, threshold)
    H, W = s.shape
    p = np.pad(s, 1)
    keep = s > 0
    for dy in (-1, 0, 1):
        for dx in (-1, 0, 1):
            if dx == 0 and dy == 0:
                continue
            keep &= s > p[1 + dy:1 + dy + H, 1 + dx:1 + dx + W]
    # cv::FAST only emits rows 3..H-4 / cols 3..W-4; scores outside are 0 anyway
    ys, xs = np.nonzero(keep)
    return xs.astype(np.int32), ys.astype(np.int32), s[ys, xs]


def retain_best(values, n):
    """KeyPointsFilter::retainBest: indices kept (all with value >= the n-th largest)."""
    if n < 0 or len(values) <= n:
        return np.arange(len(values))
    if n == 0:
        return np.zeros(0, np.int64)
    thr = np.sort(values)[::-1][n - 1]
    return np.nonzero(values >= thr)[0]


# ---- B.5: Harris response (7x7 block) ----------------------------------------------------------------
def harris_responses(ext, border, xs, ys, block=7):
    I = ext.astype(np.int32)
    r = block // 2
    scale = F32(1.0) / (F32(4 * block) * F32(255.0))
    scale4 = F32(F32(F32(scale * scale) * scale) * scale)  # scale*scale*scale*scale, left to right
    out = np.zeros(len(xs), F32)
    for k, (x, y) in enumerate(zip(xs, ys)):
        cx, cy = x + border, y + border
        win = I[cy - r - 1:cy + r + 2, cx - r - 1:cx + r + 2]
        Ix = (win[1:-1, 2:] - win[1:-1, :-2]) * 2 + (win[:-2, 2:] - win[:-2, :-2]) + (win[2:, 2:] - win[2:, :-2])
        Iy = (win[2:, 1:-1] - win[:-2, 1:-1]) * 2 + (win[2:, :-2] - win[:-2, :-2]) + (win[2:, 2:] - win[:-2, 2:])
        a = int((Ix * Ix).sum())
        b = int((Iy * Iy).sum())
        c = int((Ix * Iy).sum())
        fa, fb, fc = F32(a), F32(b), F32(c)
        # ((float)a*b - (float)c*c - harris_k*((float)a+b)*((float)a+b)) * scale^4, C precedence, no FMA
        out[k] = F32(F32(F32(fa * fb) - F32(fc * fc)) - F32(F32(HARRIS_K * F32(fa + fb)) * F32(fa + fb))) * scale4
    return out


# ---- B.7: intensity-centroid angle with cv::fastAtan2 ----------------------------------------------
UMAX = [15, 15, 15, 15, 14, 14, 14, 13, 13, 12, 11, 10, 9, 8, 6, 3]


def fast_atan2(y, x):
    y = np.asarray(y, F32)
    x = np.asarray(x, F32)
    scale = F32(180.0 / np.pi)
    p1 = F32(0.9997878412794807) * scale
    p3 = F32(-0.3258083974640975) * scale
    p5 = F32(0.1555786518463281) * scale
    p7 = F32(-0.04432655554792128) * scale
    ax, ay = np.abs(x), np.abs(y)
    eps = F32(2.2204460492503131e-16)
    swap = ax < ay
    mn = np.where(swap, ax, ay)
    mx = np.where(swap, ay, ax)
    c = (mn / (mx + eps)).astype(F32)
    c2 = (c * c).astype(F32)
    a = ((((p7 * c2).astype(F32) + p5).astype(F32) * c2).astype(F32) + p3).astype(F32)
    a = (((a * c2).astype(F32) + p1).astype(F32) * c).astype(F32)
    a = np.where(swap, (F32(90) - a).astype(F32), a)
    a = np.where(x < 0, (F32(180) - a).astype(F32), a)
    a = np.where(y < 0, (F32(360) - a).astype(F32), a)
    return a.astype(F32)


def ic_angles(ext, border, xs, ys):
    I = ext.astype(np.int64)
    out = np.zeros(len(xs), F32)
    m01s = np.zeros(len(xs), np.int64)
    m10s = np.zeros(len(xs), np.int64)
    for k, (x, y) in enumerate(zip(xs, ys)):
        cx, cy = x + border, y + border
        m01 = 0
        m10 = 0
        u = np.arange(-HALF_PATCH, HALF_PATCH + 1)
        m10 += int((u * I[cy, cx - HALF_PATCH:cx + HALF_PATCH + 1]).sum())
        for v in range(1, HALF_PATCH + 1):
            d = UMAX[v]
            uu = np.arange(-d, d + 1)
            below = I[cy + v, cx - d:cx + d + 1]
            above = I[cy - v, cx - d:cx + d + 1]
            m10 += int((uu * (below + above)).sum())
            m01 += v * int((below - above).sum())
        m01s[k], m10s[k] = m01, m10
    out = fast_atan2(m01s.astype(F32), m10s.astype(F32))
    return out


# ---- B.6: 7x7 sigma=2 float separable blur with OpenCV's FMA placement ---------------------------------
def gaussian_kernel7():
    import cv2
    return cv2.getGaussianKernel(7, 2, cv2.CV_32F).reshape(-1).astype(F32)


def _fma(a, b, c):
    # emulate a single-rounding float fma(a, b, c) through float64 (exact product of two floats fits in
    # float64; the sum is then rounded once to float64 and once to float32 -- double rounding can differ
    # from a true fma only when the float64 result lies exactly on a float32 tie, which cannot happen here
    # because a*b+c carries at most 48+ significant bits that are checked in tests against cv2)
    return (a.astype(np.float64) * b.astype(np.float64) + c.astype(np.float64)).astype(F32)


def blur7_level(ext, border, rows, cols, k=None):
    """Blur of the level ROI inside its reflect-101 bordered image (what ORB does before descriptors)."""
    if k is None:
        k = gaussian_kernel7()
    S = ext.astype(F32)
    # row pass over the rows the column pass needs: y in [-3, rows+3), x in [0, cols)
    y0 = border - 3
    R = np.zeros((rows + 6, cols), F32)
    base = S[y0:y0 + rows + 6, :]
    acc = (k[0] * base[:, border - 3:border - 3 + cols]).astype(F32)
    for j in range(1, 7):
        acc = _fma(np.full_like(acc, k[j]), base[:, border - 3 + j:border - 3 + j + cols], acc)
    R = acc
    out = (k[3] * R[3:3 + rows]).astype(F32)
    for j in range(1, 4):
        pair = (R[3 + j:3 + j + rows] + R[3 - j:3 - j + rows]).astype(F32)
        out = _fma(np.full_like(out, k[3 + j]), pair, out)
    return np.clip(np.rint(out), 0, 255).astype(np.uint8)


# ---- B.8: rBRIEF-256 ---------------------------------------------------------------------------------
def rbrief(blur_ext, border, xs, ys, angles_deg, pattern=None):
    if pattern is None:
        pattern = bit_pattern_31()
    pts = pattern.reshape(-1, 2).astype(F32)  # 512 points (x, y)
    out = np.zeros((len(xs), 32), np.uint8)
    for k, (x, y, ang) in enumerate(zip(xs, ys, angles_deg)):
        ar = F32(F32(ang) * F32(np.pi / 180.0))
        a = F32(np.cos(np.float64(ar)))
        b = F32(np.sin(np.float64(ar)))
        px, py = pts[:, 0], pts[:, 1]
        xx = (px * a).astype(F32) - (py * b).astype(F32)
        yy = (px * b).astype(F32) + (py * a).astype(F32)
        ix = np.rint(xx.astype(F32)).astype(np.int64)
        iy = np.rint(yy.astype(F32)).astype(np.int64)
        vals = blur_ext[y + border + iy, x + border + ix].astype(np.int32)
        bits = (vals[0::2] < vals[1::2]).astype(np.uint8)
        out[k] = np.packbits(bits.reshape(32, 8)[:, ::-1], axis=1).reshape(-1)
    return out


# ---- the whole extractor ---------------------------------------------------------------------------------
def orb_detect_and_compute(img, nfeatures=2000, nlevels=8, scale_factor=1.2, fast_threshold=20, return_levels=False):
    """Returns dict of arrays: x, y (float32 image coords), size, angle, response, octave, desc (N,32);
    keypoints are ordered by (octave, y_level, x_level) -- cv2's in-level order is implementation defined."""
    levels, scales = build_pyramid(img, nlevels, scale_factor)
    quotas = level_quotas(nfeatures, nlevels, scale_factor)
    pattern = bit_pattern_31()
    k7 = gaussian_kernel7()
    res = {k: [] for k in ("x", "y", "size", "angle", "response", "octave", "desc", "lx", "ly")}
    for l, lv in enumerate(levels):
        H, W = lv.shape
        xs, ys, sc = fast9_detect(lv, fast_threshold)
        inb = (xs >= EDGE) & (xs < W - EDGE) & (ys >= EDGE) & (ys < H - EDGE)
        xs, ys, sc = xs[inb], ys[inb], sc[inb]
        keep = retain_best(sc.astype(np.float32), 2 * quotas[l])
        xs, ys = xs[keep], ys[keep]
        ext = reflect101(lv, 32)
        hr = harris_responses(ext, 32, xs, ys)
        keep = retain_best(hr, quotas[l])
        xs, ys, hr = xs[keep], ys[keep], hr[keep]
        order = np.lexsort((xs, ys))
        xs, ys, hr = xs[order], ys[order], hr[order]
        ang = ic_angles(ext, 32, xs, ys)
        bl = blur7_level(ext, 32, H, W, k7)
        bl_ext = ext.copy()
        bl_ext[32:32 + H, 32:32 + W] = bl
        # the pyramid image is blurred in place inside its border: samples never leave the level ROI for
        # keypoints >= 31 px from the edge (pattern reach <= 13*sqrt(2) < 19)
        desc = rbrief(bl_ext, 32, xs, ys, ang, pattern)
        s = scales[l]
        res["x"].append((xs.astype(F32) * s).astype(F32))
        res["y"].append((ys.astype(F32) * s).astype(F32))
        res["size"].append(np.full(len(xs), F32(PATCH) * s, F32))
        res["angle"].append(ang)
        res["response"].append(hr)
        res["octave"].append(np.full(len(xs), l, np.int32))
        res["desc"].append(desc)
        res["lx"].append(xs)
        res["ly"].append(ys)
    out = {k: (np.concatenate(v) if len(v) else np.zeros(0)) for k, v in res.items()}
    if return_levels:
        out["levels"] = levels
    return out


# ---- B.10: BFMatcher(NORM_HAMMING).knnMatch(k=2) -----------------------------------------------------------
_POP = np.array([bin(i).count("1") for i in range(256)], np.int32)


def knn2(d1, d2, chunk=256):
    n1 = len(d1)
    idx = np.zeros((n1, 2), np.int32)
    dist = np.zeros((n1, 2), np.int32)
    for s in range(0, n1, chunk):
        D = _POP[d1[s:s + chunk, None, :] ^ d2[None, :, :]].sum(-1)
        order = np.lexsort((np.broadcast_to(np.arange(len(d2)), D.shape), D), axis=1)[:, :2]
        idx[s:s + chunk] = order
        dist[s:s + chunk] = np.take_along_axis(D, order, 1)
    return idx, dist
