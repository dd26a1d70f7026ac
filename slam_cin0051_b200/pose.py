"""Host-side mirror of slam::PoseEstimator (include/slam/frontend/pose_estimator.hpp:13-36) up to the essential
matrix: PoseEstimator::estimate gathers the matched keypoint coordinates (pose_estimator.cpp:30-35) and calls
cv::findEssentialMat(points1, points2, K, cv::RANSAC) (:42) -- here slamcu_find_essential on the device."""
from __future__ import annotations

import ctypes as C

import numpy as np

from ._lib import Context
from .common import Camera


def find_essential(points1, points2, K4, prob: float = 0.999, threshold: float = 1.0, max_iters: int = 1000,
                   context: Context | None = None):
    """cv::findEssentialMat(p1, p2, K, RANSAC, prob, threshold, maxIters).  Returns (E 3x3 float64 or None, mask uint8[n],
    n_inliers).  Raises ValueError below 6 correspondences."""
    ctx = context or Context.default()
    p1 = np.ascontiguousarray(points1, np.float32).reshape(-1, 2)
    p2 = np.ascontiguousarray(points2, np.float32).reshape(-1, 2)
    if len(p1) != len(p2):
        raise RuntimeError("point sets must have the same size")
    k = np.ascontiguousarray(K4, np.float64)
    E = np.zeros(9, np.float64)
    mask = np.zeros(max(len(p1), 1), np.uint8)
    n_in = C.c_int(0)
    ctx.check(ctx.lib.slamcu_find_essential(ctx.handle, p1.ctypes.data, p2.ctypes.data, len(p1), k.ctypes.data, prob, threshold,
                                            max_iters, E.ctypes.data, mask.ctypes.data, C.byref(n_in)))
    mask = mask[: len(p1)]
    if n_in.value <= 0:
        return None, mask, 0
    return E.reshape(3, 3), mask, n_in.value


def estimate_pose(points1, points2, K4, context: Context | None = None):
    """PoseEstimator::estimate on gathered pixel correspondences: findEssentialMat(RANSAC) + simpleRecoverPose.
    Returns dict(E, mask, inliers, R, t, front) or None where the reference returns early (< 8 matches / no E)."""
    ctx = context or Context.default()
    p1 = np.ascontiguousarray(points1, np.float32).reshape(-1, 2)
    p2 = np.ascontiguousarray(points2, np.float32).reshape(-1, 2)
    if len(p1) < 8:
        return None
    k = np.ascontiguousarray(K4, np.float64)
    E, R, t = np.zeros(9), np.zeros(9), np.zeros(3)
    mask = np.zeros(len(p1), np.uint8)
    front = np.zeros(4, np.int32)
    n_in = C.c_int(0)
    st = ctx.lib.slamcu_estimate_pose(ctx.handle, p1.ctypes.data, p2.ctypes.data, len(p1), k.ctypes.data, E.ctypes.data,
                                      mask.ctypes.data, C.byref(n_in), R.ctypes.data, t.ctypes.data, front.ctypes.data)
    if st == 2:  # SLAMCU_EMPTY_INPUT: the reference's early returns
        return None
    ctx.check(st)
    return {"E": E.reshape(3, 3), "mask": mask, "inliers": n_in.value, "R": R.reshape(3, 3), "t": t.reshape(3, 1), "front": front}


def fivept_solve(x1, x2, context: Context | None = None):
    """5-point minimal solver probe: x1, x2 (S, 5, 2) normalised points -> list of (k_s, 3, 3) arrays."""
    ctx = context or Context.default()
    a = np.ascontiguousarray(x1, np.float64).reshape(-1, 5, 2)
    b = np.ascontiguousarray(x2, np.float64).reshape(-1, 5, 2)
    S = len(a)
    models = np.zeros((S, 10, 9), np.float64)
    counts = np.zeros(S, np.int32)
    ctx.check(ctx.lib.slamcu_fivept_solve(ctx.handle, a.ctypes.data, b.ctypes.data, S, models.ctypes.data, counts.ctypes.data))
    return [models[s, : counts[s]].reshape(-1, 3, 3).copy() for s in range(S)]


def ransac_score(models, x1, x2, thr2: float, with_masks: bool = False, context: Context | None = None):
    """Sampson-error inlier counts of M essential-matrix hypotheses against n normalised correspondences."""
    ctx = context or Context.default()
    m = np.ascontiguousarray(models, np.float64).reshape(-1, 9)
    a = np.ascontiguousarray(x1, np.float64).reshape(-1, 2)
    b = np.ascontiguousarray(x2, np.float64).reshape(-1, 2)
    counts = np.zeros(len(m), np.int32)
    masks = np.zeros((len(m), len(a)), np.uint8) if with_masks else None
    ctx.check(ctx.lib.slamcu_ransac_score(ctx.handle, m.ctypes.data, len(m), a.ctypes.data, b.ctypes.data, len(a), thr2,
                                          counts.ctypes.data, masks.ctypes.data if with_masks else None))
    return (counts, masks) if with_masks else counts


class PoseEstimator:
    """slam::PoseEstimator up to E and the inlier mask.  `estimate` mirrors the reference's guard: fewer than 8 matches
    -> returns None without touching anything (pose_estimator.cpp:22-26)."""

    def __init__(self, camera: Camera, context: Context | None = None):
        self.camera = camera
        self.ctx = context or camera.ctx

    def estimate_essential(self, keypoints1, keypoints2, matches):
        """keypoints: KEYPOINT_DTYPE arrays; matches: MATCH_DTYPE array or (n, 2) index pairs."""
        m = np.asarray(matches)
        if m.dtype.names:
            q, t = m["queryIdx"], m["trainIdx"]
        else:
            m = m.reshape(-1, 2)
            q, t = m[:, 0], m[:, 1]
        if len(q) < 8:
            return None
        p1 = np.stack([keypoints1["x"][q], keypoints1["y"][q]], 1)
        p2 = np.stack([keypoints2["x"][t], keypoints2["y"][t]], 1)
        c = self.camera
        return find_essential(p1, p2, (c.fx, c.fy, c.cx, c.cy), context=self.ctx)

    def estimate(self, keypoints1, keypoints2, matches):
        """PoseEstimator::estimate: returns (R 3x3, t 3x1) or None where the reference leaves R, t untouched."""
        m = np.asarray(matches)
        if m.dtype.names:
            q, t = m["queryIdx"], m["trainIdx"]
        else:
            m = m.reshape(-1, 2)
            q, t = m[:, 0], m[:, 1]
        if len(q) < 8:
            return None
        p1 = np.stack([keypoints1["x"][q], keypoints1["y"][q]], 1)
        p2 = np.stack([keypoints2["x"][t], keypoints2["y"][t]], 1)
        c = self.camera
        r = estimate_pose(p1, p2, (c.fx, c.fy, c.cx, c.cy), context=self.ctx)
        return None if r is None else (r["R"], r["t"])


def triangulate(P1, P2, points1, points2, context: Context | None = None):
    """slam::triangulate (common.hpp:201-221): per correspondence the null vector of the 4x4 DLT system of the 3x4 projection
    matrices P1, P2 and the pixel coordinates.  Returns (points4 (n, 4) unit norm with w >= 0, points3 (n, 3) = x / w)."""
    ctx = context or Context.default()
    a = np.ascontiguousarray(P1, np.float64).reshape(12)
    b = np.ascontiguousarray(P2, np.float64).reshape(12)
    p1 = np.ascontiguousarray(points1, np.float32).reshape(-1, 2)
    p2 = np.ascontiguousarray(points2, np.float32).reshape(-1, 2)
    if len(p1) != len(p2):
        raise RuntimeError("point sets must have the same size")
    x4, x3 = np.zeros((len(p1), 4)), np.zeros((len(p1), 3))
    ctx.check(ctx.lib.slamcu_triangulate(ctx.handle, a.ctypes.data, b.ctypes.data, p1.ctypes.data, p2.ctypes.data, len(p1), x4.ctypes.data,
                                         x3.ctypes.data))
    return x4, x3


def triangulate_points(K, R, t, points1, points2, context: Context | None = None):
    """PoseEstimator::triangulatePoints (pose_estimator.cpp:69-104) on gathered pixel correspondences: P1 = K [I | 0],
    P2 = K [R | t]; returns the (n, 3) points."""
    K = np.asarray(K, np.float64).reshape(3, 3)
    T1 = np.hstack([np.eye(3), np.zeros((3, 1))])
    T2 = np.hstack([np.asarray(R, np.float64).reshape(3, 3), np.asarray(t, np.float64).reshape(3, 1)])
    return triangulate(K @ T1, K @ T2, points1, points2, context)[1]


def pnp_sample_indices(seed: int, n: int, iterations: int) -> np.ndarray:
    """The 6-subset draws of LoopClosure::verifyGeometricConsistency (loop_closure.cpp:177-193), libstdc++ mt19937 seeded by `seed`."""
    from . import _lib
    out = np.zeros((iterations, 6), np.int32)
    if _lib.load().slamcu_pnp_sample_indices(seed & 0xFFFFFFFF, n, iterations, out.ctypes.data) != 0:
        raise RuntimeError("pnp_sample_indices: need n >= 6")
    return out


def pnp_ransac(points3d, points2d, K, samples6, threshold: float = 2.0, context: Context | None = None):
    """The hypothesis loop of LoopClosure::verifyGeometricConsistency (loop_closure.cpp:177-222) for the given 6-subsets.
    Returns (counts (h, 2) int32, Rt (h, 2, 12)): inliers, R (row-major) and t for both signs of each DLT null vector."""
    ctx = context or Context.default()
    X = np.ascontiguousarray(points3d, np.float64).reshape(-1, 3)
    x = np.ascontiguousarray(points2d, np.float64).reshape(-1, 2)
    s6 = np.ascontiguousarray(samples6, np.int32).reshape(-1, 6)
    k = np.ascontiguousarray(K, np.float64).reshape(9)
    counts = np.zeros((len(s6), 2), np.int32)
    Rt = np.zeros((len(s6), 2, 12))
    ctx.check(ctx.lib.slamcu_pnp_ransac(ctx.handle, X.ctypes.data, x.ctypes.data, len(X), s6.ctypes.data, len(s6), k.ctypes.data, threshold,
                                        counts.ctypes.data, Rt.ctypes.data))
    return counts, Rt


def verify_geometric_consistency(points3d, points2d, K, iterations: int = 100, threshold: float = 2.0, min_inliers: int = 5, seed: int = 0,
                                 sign: str = "best", context: Context | None = None):
    """LoopClosure::verifyGeometricConsistency's decision (loop_closure.cpp:177-236) on gathered correspondences: the first
    hypothesis with the strictly largest inlier count wins (:217-221); None below `min_inliers` (:224-235).  sign: which null-vector
    sign the SVD 'returned' ("best": the better of the two per hypothesis; 0 / 1: fixed)."""
    s6 = pnp_sample_indices(seed, len(points3d), iterations)
    counts, Rt = pnp_ransac(points3d, points2d, K, s6, threshold, context)
    pick = counts.argmax(1) if sign == "best" else np.full(len(counts), int(sign))
    c = counts[np.arange(len(counts)), pick]
    best, max_inl = -1, 0
    for h in range(len(c)):
        if c[h] > max_inl:
            best, max_inl = h, int(c[h])
    if best < 0 or max_inl < min_inliers:
        return None
    rt = Rt[best, pick[best]]
    T = np.eye(4)
    T[:3, :3] = rt[:9].reshape(3, 3)
    T[:3, 3] = rt[9:]
    return {"inliers": max_inl, "relativeTransform": T, "hypothesis": best}
