"""CPU restatement of the geometric-verification RANSAC of LoopClosure (reference: src/backend/loop_closure.cpp:177-274) and of
slam::triangulate (include/slam/common/common.hpp:201-221) -- TEST INFRASTRUCTURE ONLY (numpy; only tests/ may import it).

Parity unpinned by the reference: it holds no golden values for these paths, Eigen (JacobiSVD) is not in this image, and both
functions are defective as written: solvePnP reads its row-major DLT vector back through a column-major
Eigen::Map and never removes K; triangulate's copyTo does not write its output.  This module restates solvePnP LITERALLY (defects
included), with numpy's SVD standing in for Eigen's and both signs of the null vector evaluated, and restates what
triangulate evidently means.
"""
from __future__ import annotations

import numpy as np


def solve_pnp_literal(X6, x6, sign_index):
    """loop_closure.cpp:238-274 for six correspondences; sign_index 0: the null vector's largest component positive, 1: negated."""
    A = np.zeros((12, 12))
    for i in range(6):
        X, Y, Z = X6[i]
        u, v = x6[i]
        A[2 * i] = [X, Y, Z, 1, 0, 0, 0, 0, -u * X, -u * Y, -u * Z, -u]          # :250
        A[2 * i + 1] = [0, 0, 0, 0, X, Y, Z, 1, -v * X, -v * Y, -v * Z, -v]      # :251
    p = np.linalg.svd(A)[2][-1]                                                    # :254-255 matrixV().col(11)
    if p[np.abs(p).argmax()] < 0:
        p = -p
    if sign_index:
        p = -p
    P = p.reshape(4, 3).T                                                          # :258 Map<Matrix<double,3,4>> is column-major
    R, t = P[:, :3], P[:, 3]                                                       # :260-261
    U, _, Vt = np.linalg.svd(R)                                                    # :263
    det = np.linalg.det(U @ Vt)                                                    # :264
    rotation = U @ np.diag([1.0, 1.0, det]) @ Vt                                   # :265-268
    translation = t / np.linalg.norm(R)                                            # :269 (Frobenius norm)
    return rotation, translation


def score(rotation, translation, X, x, K, threshold):
    """loop_closure.cpp:201-215."""
    T = X @ rotation.T + translation
    front = T[:, 2] > 0
    with np.errstate(divide="ignore", invalid="ignore"):
        proj = (T / T[:, 2:3]) @ np.asarray(K, np.float64).reshape(3, 3).T
        err = np.linalg.norm(x - proj[:, :2], axis=1)
    return int((front & (err < threshold)).sum())


def pnp_ransac(X, x, K, samples6, threshold):
    """Returns (counts (h, 2), Rt (h, 2, 12)) like slamcu_pnp_ransac."""
    X = np.asarray(X, np.float64).reshape(-1, 3)
    x = np.asarray(x, np.float64).reshape(-1, 2)
    counts = np.zeros((len(samples6), 2), np.int32)
    Rt = np.zeros((len(samples6), 2, 12))
    for h, idx in enumerate(samples6):
        for s in range(2):
            R, t = solve_pnp_literal(X[idx], x[idx], s)
            counts[h, s] = score(R, t, X, x, K, threshold)
            Rt[h, s, :9] = R.reshape(9)
            Rt[h, s, 9:] = t
    return counts, Rt


def triangulate(P1, P2, pts1, pts2):
    """common.hpp:201-221 (A rows :208-211, null vector :213-215); returns (x4 unit norm with w >= 0, x3 = x / w)."""
    P1 = np.asarray(P1, np.float64).reshape(3, 4)
    P2 = np.asarray(P2, np.float64).reshape(3, 4)
    x4 = np.zeros((len(pts1), 4))
    for i, (a, b) in enumerate(zip(np.asarray(pts1, np.float32), np.asarray(pts2, np.float32))):
        A = np.stack([float(a[0]) * P1[2] - P1[0], float(a[1]) * P1[2] - P1[1], float(b[0]) * P2[2] - P2[0], float(b[1]) * P2[2] - P2[1]])
        v = np.linalg.svd(A)[2][-1]
        x4[i] = -v if v[3] < 0 else v
    return x4, x4[:, :3] / x4[:, 3:4]
