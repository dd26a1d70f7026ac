"""CPU restatement of cv::findEssentialMat(p1, p2, K, RANSAC, prob, threshold, maxIters) -- TEST INFRASTRUCTURE ONLY.

The reference calls it at src/frontend/pose_estimator.cpp:42 with the defaults (0.999, 1.0, 1000).  The code lives in
OpenCV (calib3d five-point.cpp / ptsetreg.cpp; conanfile.txt:2 pins opencv/4.12.0, not vendored), so this module
restates the published algorithm and is pinned against cv2 itself in
tests/test_essential_oracle.py:
  * RNG (cv::RNG multiply-with-carry, state 0xFFFFFFFFFFFFFFFF), 5-subset sampling, Sampson-error inlier test in
    double narrowed to float, the accept rule and RANSACUpdateNumIters: reproduced exactly;
  * the 5-point minimal solver (Nister): same mathematics -- null space of the 5x9 epipolar system, the ten cubic
    constraints det(E) = 0 and 2 E E'E - tr(E E')E = 0, Gauss-Jordan on the 10x20 coefficient matrix, a degree-10
    polynomial in z, real roots, back-substitution -- but with its own null-space basis and root finder, so the
    candidate E's agree with cv2's to rounding (compared up to sign) and their order within one sample may differ.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this module.
"""
from __future__ import annotations

import numpy as np

F32 = np.float32
M32 = 0xFFFFFFFF


class CvRNG:
    """cv::RNG: state = (uint32)state * 4164903690 + (state >> 32); returns (uint32)state."""

    def __init__(self, state=0xFFFFFFFFFFFFFFFF):
        self.state = state

    def next(self):
        self.state = ((self.state & M32) * 4164903690 + (self.state >> 32)) & 0xFFFFFFFFFFFFFFFF
        return self.state & M32

    def uniform(self, a, b):
        return a if a == b else self.next() % (b - a) + a


def sample_indices(rng, count, m=5):
    """RANSACPointSetRegistrator::getSubset: m distinct indices, redraw on duplicates."""
    idx = []
    while len(idx) < m:
        v = rng.uniform(0, count)
        if v in idx:
            continue
        idx.append(v)
    return idx


# ---- polynomial helpers: monomials of degree <= 3 in (x, y, z) -------------------------------------------------------
# order of the 20 cubic monomials: the ten that are eliminated first, then the ten kept (Nister / Stewenius order)
MONO3 = [(3, 0, 0), (0, 3, 0), (2, 1, 0), (1, 2, 0), (2, 0, 1), (2, 0, 0), (0, 2, 1), (0, 2, 0), (1, 1, 1), (1, 1, 0),
         (1, 0, 2), (1, 0, 1), (1, 0, 0), (0, 1, 2), (0, 1, 1), (0, 1, 0), (0, 0, 3), (0, 0, 2), (0, 0, 1), (0, 0, 0)]
MONO3_INDEX = {m: i for i, m in enumerate(MONO3)}
MONO1 = [(1, 0, 0), (0, 1, 0), (0, 0, 1), (0, 0, 0)]
MONO2 = [(2, 0, 0), (0, 2, 0), (0, 0, 2), (1, 1, 0), (1, 0, 1), (0, 1, 1), (1, 0, 0), (0, 1, 0), (0, 0, 1), (0, 0, 0)]
MONO2_INDEX = {m: i for i, m in enumerate(MONO2)}


def _mul11(a, b):
    out = np.zeros(10)
    for i, mi in enumerate(MONO1):
        for j, mj in enumerate(MONO1):
            out[MONO2_INDEX[(mi[0] + mj[0], mi[1] + mj[1], mi[2] + mj[2])]] += a[i] * b[j]
    return out


def _mul21(a, b):
    out = np.zeros(20)
    for i, mi in enumerate(MONO2):
        for j, mj in enumerate(MONO1):
            out[MONO3_INDEX[(mi[0] + mj[0], mi[1] + mj[1], mi[2] + mj[2])]] += a[i] * b[j]
    return out


def null_space_basis(Q):
    """Orthonormal basis (4 x 9) of the null space of the 5 x 9 matrix Q: Gauss-Jordan with full pivoting, then
    modified Gram-Schmidt (twice)."""
    A = Q.astype(np.float64).copy()
    cols = list(range(9))
    for r in range(5):
        sub = np.abs(A[r:, r:])
        pr, pc = np.unravel_index(np.argmax(sub), sub.shape)
        pr += r
        pc += r
        A[[r, pr]] = A[[pr, r]]
        A[:, [r, pc]] = A[:, [pc, r]]
        cols[r], cols[pc] = cols[pc], cols[r]
        A[r] = A[r] / A[r, r]
        for k in range(5):
            if k != r:
                A[k] = A[k] - A[k, r] * A[r]
    basis = np.zeros((4, 9))
    for k in range(4):
        v = np.zeros(9)
        for r in range(5):
            v[cols[r]] = -A[r, 5 + k]
        v[cols[5 + k]] = 1.0
        basis[k] = v
    for _ in range(2):
        for k in range(4):
            for j in range(k):
                basis[k] = basis[k] - np.dot(basis[k], basis[j]) * basis[j]
            basis[k] = basis[k] / np.sqrt(np.dot(basis[k], basis[k]))
    return basis


def constraint_matrix(basis):
    """10 x 20 coefficients of det(E) and 2 E E'E - tr(E E')E for E = x B0 + y B1 + z B2 + B3 (columns: MONO3)."""
    E = [[np.array([basis[0][3 * i + j], basis[1][3 * i + j], basis[2][3 * i + j], basis[3][3 * i + j]]) for j in range(3)]
         for i in range(3)]
    A = np.zeros((10, 20))
    # det(E)
    A[0] = (_mul21(_mul11(E[0][1], E[1][2]) - _mul11(E[0][2], E[1][1]), E[2][0]) +
            _mul21(_mul11(E[0][2], E[1][0]) - _mul11(E[0][0], E[1][2]), E[2][1]) +
            _mul21(_mul11(E[0][0], E[1][1]) - _mul11(E[0][1], E[1][0]), E[2][2]))
    EEt = [[sum(_mul11(E[i][k], E[j][k]) for k in range(3)) for j in range(3)] for i in range(3)]
    tr = EEt[0][0] + EEt[1][1] + EEt[2][2]
    L = [[EEt[i][j] - (0.5 * tr if i == j else 0.0) for j in range(3)] for i in range(3)]  # E E' - tr/2 I
    r = 1
    for i in range(3):
        for j in range(3):
            A[r] = sum(_mul21(L[i][k], E[k][j]) for k in range(3))
            r += 1
    return A


def _polymul(a, b):
    return np.convolve(a, b)  # highest power first


def real_roots(c, im_tol=1e-10, max_iter=500):
    """Durand-Kerner on a real polynomial (coefficients highest power first); returns the real roots."""
    c = np.trim_zeros(np.asarray(c, np.float64), "f")
    n = len(c) - 1
    if n < 1:
        return []
    a = c / c[0]
    radius = 1.0 + np.max(np.abs(a[1:]))
    roots = np.array([0.4 + 0.9j]) ** np.arange(n) * min(radius, 1e3) ** 0  # classic start (0.4+0.9i)^k
    for _ in range(max_iter):
        delta = 0.0
        for i in range(n):
            p = roots[i]
            num = np.polyval(a, p)
            den = 1.0 + 0j
            for j in range(n):
                if j != i:
                    den *= p - roots[j]
            if den == 0:
                continue
            d = num / den
            roots[i] = p - d
            delta = max(delta, abs(d))
        if delta < 1e-15 * max(1.0, np.max(np.abs(roots))):
            break
    out = []
    for r in roots:
        # polish the real candidates with two Newton steps on the real axis
        if abs(r.imag) <= 1e-6 * max(1.0, abs(r.real)):
            x = r.real
            for _ in range(3):
                f = np.polyval(a, x)
                df = np.polyval(np.polyder(a), x)
                if df != 0:
                    x -= f / df
            if abs(r.imag) <= im_tol * max(1.0, abs(r.real)) or abs(np.polyval(a, x)) < 1e-9 * (1 + abs(np.polyval(np.abs(a), abs(x)))):
                out.append(x)
    return out


def five_point(x1, x2):
    """x1, x2: (5, 2) normalised image points.  Returns a list of 3x3 essential matrices with unit Frobenius norm."""
    x1 = np.asarray(x1, np.float64)
    x2 = np.asarray(x2, np.float64)
    Q = np.stack([x2[:, 0] * x1[:, 0], x2[:, 1] * x1[:, 0], x1[:, 0], x2[:, 0] * x1[:, 1], x2[:, 1] * x1[:, 1], x1[:, 1],
                  x2[:, 0], x2[:, 1], np.ones(5)], 1)
    # OpenCV's column order is (x2x1, y2x1, x1, x2y1, y2y1, y1, x2, y2, 1), i.e. the vector is E transposed-major:
    # x2' E x1 with e = (E00, E10, E20, E01, E11, E21, E02, E12, E22); reorder to row-major E for the algebra
    perm = [0, 3, 6, 1, 4, 7, 2, 5, 8]
    Qr = np.zeros_like(Q)
    Qr[:, perm] = Q
    basis = null_space_basis(Qr)  # rows: row-major 3x3 matrices
    A = constraint_matrix(basis)
    # Gauss-Jordan on the first ten columns, partial pivoting
    M = A.copy()
    for col in range(10):
        piv = col + int(np.argmax(np.abs(M[col:, col])))
        if M[piv, col] == 0:
            return []
        M[[col, piv]] = M[[piv, col]]
        M[col] = M[col] / M[col, col]
        for r in range(10):
            if r != col:
                M[r] = M[r] - M[r, col] * M[col]
    R = M[:, 10:]
    # rows 4..9: x^2 z, x^2, y^2 z, y^2, xyz, xy ; kept monomials: xz^2 xz x yz^2 yz y z^3 z^2 z 1
    B = []
    for i in range(3):
        a, b = R[4 + 2 * i], R[5 + 2 * i]
        px = np.array([0.0, a[0], a[1], a[2]]) - np.array([b[0], b[1], b[2], 0.0])
        py = np.array([0.0, a[3], a[4], a[5]]) - np.array([b[3], b[4], b[5], 0.0])
        p1 = np.array([0.0, a[6], a[7], a[8], a[9]]) - np.array([b[6], b[7], b[8], b[9], 0.0])
        B.append((px, py, p1))
    det = (_polymul(_polymul(B[0][0], B[1][1]) - _polymul(B[0][1], B[1][0]), B[2][2]) +
           _polymul(_polymul(B[0][1], B[1][2][1:] if False else B[1][2]), B[2][0])[0:0].sum() * 0)
    # explicit cofactor expansion along the third column (degrees: px, py 3; p1 4 -> total 10)
    m01 = _polymul(B[1][0], B[2][1]) - _polymul(B[1][1], B[2][0])
    m02 = _polymul(B[0][0], B[2][1]) - _polymul(B[0][1], B[2][0])
    m03 = _polymul(B[0][0], B[1][1]) - _polymul(B[0][1], B[1][0])
    det = _polymul(B[0][2], m01) - _polymul(B[1][2], m02) + _polymul(B[2][2], m03)
    out = []
    for z in real_roots(det):
        zp = np.array([z ** 3, z ** 2, z, 1.0])
        zq = np.array([z ** 4, z ** 3, z ** 2, z, 1.0])
        Bz = np.array([[np.dot(B[i][0], zp), np.dot(B[i][1], zp), np.dot(B[i][2], zq)] for i in range(3)])
        # null vector of the 3x3 (rank 2) matrix: the largest cross product of two rows
        cr = [np.cross(Bz[0], Bz[1]), np.cross(Bz[0], Bz[2]), np.cross(Bz[1], Bz[2])]
        v = max(cr, key=lambda c: np.dot(c, c))
        if abs(v[2]) < 1e-10 * np.sqrt(np.dot(v, v)):
            continue
        x, y = v[0] / v[2], v[1] / v[2]
        e = x * basis[0] + y * basis[1] + z * basis[2] + basis[3]
        e = e / np.sqrt(np.dot(e, e))
        out.append(e.reshape(3, 3))
    return out


def sampson_errors(E, x1, x2):
    """EMEstimatorCallback::computeError: float(err) per point, double arithmetic in OpenCV's operation order."""
    ax, ay = x1[:, 0], x1[:, 1]
    bx, by = x2[:, 0], x2[:, 1]
    e0 = (E[0, 0] * ax + E[0, 1] * ay) + E[0, 2] * 1.0
    e1 = (E[1, 0] * ax + E[1, 1] * ay) + E[1, 2] * 1.0
    e2 = (E[2, 0] * ax + E[2, 1] * ay) + E[2, 2] * 1.0
    t0 = (E[0, 0] * bx + E[1, 0] * by) + E[2, 0] * 1.0
    t1 = (E[0, 1] * bx + E[1, 1] * by) + E[2, 1] * 1.0
    dot = (bx * e0 + by * e1) + 1.0 * e2
    return (dot * dot / (e0 * e0 + e1 * e1 + t0 * t0 + t1 * t1)).astype(F32)


def update_num_iters(p, ep, model_points, max_iters):
    p = min(max(p, 0.0), 1.0)
    ep = min(max(ep, 0.0), 1.0)
    num = max(1.0 - p, np.finfo(np.float64).tiny)
    denom = 1.0 - (1.0 - ep) ** model_points
    if denom < np.finfo(np.float64).tiny:
        return 0
    num = np.log(num)
    denom = np.log(denom)
    return max_iters if (denom >= 0 or -num >= max_iters * (-denom)) else int(np.rint(num / denom))


def normalise(p, K4):
    fx, fy, cx, cy = K4
    p = np.asarray(p, np.float32).astype(np.float64)
    return np.stack([(p[:, 0] - cx) / fx, (p[:, 1] - cy) / fy], 1)


def find_essential(p1, p2, K4, prob=0.999, threshold=1.0, max_iters=1000, solver=five_point, return_trace=False):
    """Returns (E (3x3) or None, mask uint8[n], n_inliers).  p1, p2: (n, 2) float32 pixel coordinates."""
    x1, x2 = normalise(p1, K4), normalise(p2, K4)
    n = len(x1)
    thr = threshold / ((K4[0] + K4[1]) / 2.0)
    t2 = F32(thr * thr)
    rng = CvRNG()
    niters = max_iters
    best, bestE, bestmask = 0, None, np.zeros(n, np.uint8)
    it = 0
    trace = []
    while it < niters:
        idx = sample_indices(rng, n)
        models = solver(x1[idx], x2[idx])
        for E in models:
            err = sampson_errors(E, x1, x2)
            mask = err <= t2
            good = int(mask.sum())
            trace.append((it, good))
            if good > max(best, 4):
                best, bestE, bestmask = good, E, mask.astype(np.uint8)
                niters = update_num_iters(prob, (n - good) / n, 5, niters)
        it += 1
    if return_trace:
        return bestE, bestmask, best, trace, it
    return bestE, bestmask, best


# ---- simpleRecoverPose (src/frontend/simple_pose_recover.cpp:6-97) as PoseEstimator::estimate calls it ---------------------
def _svd(a, backend):
    """(w, u, vt) with cv::SVD's conventions when backend is cv2 (cv2.SVDecomp), else LAPACK's."""
    if backend is not None:
        w, u, vt = backend.SVDecomp(np.ascontiguousarray(a, np.float64))
        return w.ravel(), u, vt
    u, w, vt = np.linalg.svd(np.asarray(a, np.float64))
    return w, u, vt


def simple_recover_pose(E, p1, p2, K4, backend=None):
    """E: 3x3; p1, p2: (n, 2) float32 PIXEL coordinates; returns (R, t (3,1), front[4]).

    Follows pose_estimator.cpp:53-66 + simple_pose_recover.cpp literally: the points are normalised by K and stored as
    float (cv::Point2f), yet the candidate projections are K [R|t]; first maximum of the cheirality vote wins."""
    fx, fy, cx, cy = K4
    K = np.array([[fx, 0, cx], [0, fy, cy], [0, 0, 1.0]])
    n1 = np.stack([((p1[:, 0].astype(np.float64) - cx) / fx), ((p1[:, 1].astype(np.float64) - cy) / fy)], 1).astype(F32)
    n2 = np.stack([((p2[:, 0].astype(np.float64) - cx) / fx), ((p2[:, 1].astype(np.float64) - cy) / fy)], 1).astype(F32)
    w, u, vt = _svd(E, backend)
    W = np.array([[0, -1, 0], [1, 0, 0], [0, 0, 1.0]])
    R1, R2, t = u @ W @ vt, u @ W.T @ vt, u[:, 2:3].copy()
    if np.linalg.det(R1) < 0:
        R1 = -R1
    if np.linalg.det(R2) < 0:
        R2 = -R2
    cands = [(R1, t), (R2, t), (R1, -t), (R2, -t)]
    KP0 = K @ np.eye(3, 4)
    front = []
    for R, tt in cands:
        KP = K @ np.hstack([R, tt])
        cnt = 0
        for a, b in zip(n1.astype(np.float64), n2.astype(np.float64)):
            A = np.stack([a[0] * KP0[2] - KP0[0], a[1] * KP0[2] - KP0[1], b[0] * KP[2] - KP[0], b[1] * KP[2] - KP[1]])
            _, _, v = _svd(A, backend)
            X = v[3] / v[3][3]
            cnt += (X[2] > 0) and ((KP @ X)[2] > 0)
        front.append(int(cnt))
    best = int(np.argmax(front))  # first maximum
    R, tt = cands[best]
    return R, tt, front
