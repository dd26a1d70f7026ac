"""SURVEY 8(f) rows 3 and 4 on the device: slam::triangulate / PoseEstimator::triangulatePoints and the geometric-verification
RANSAC of LoopClosure (solvePnP restated literally + the reprojection scoring loop), against the numpy restatements in
oracle/pnp_oracle.py.  Stated tolerance: triangulated points within 1e-7 relative of numpy's SVD solution; PnP inlier counts
identical, R and t within 1e-8."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _scene(rng, n):
    import cv2
    K = np.array([[525.0, 0, 319.5], [0, 525.0, 239.5], [0, 0, 1.0]])
    X = np.c_[rng.uniform(-3, 3, n), rng.uniform(-2, 2, n), rng.uniform(4, 12, n)]
    R, _ = cv2.Rodrigues(rng.normal(0, 0.08, 3))
    t = rng.normal(0, 0.4, 3)
    a = (K @ X.T).T
    b = (K @ (X @ R.T + t).T).T
    return K, X, R, t, (a[:, :2] / a[:, 2:]).astype(np.float32), (b[:, :2] / b[:, 2:]).astype(np.float32)


def test_triangulate_equals_numpy_and_recovers_the_scene(gpu_ctx):
    import slam_cin0051_b200.pose as P
    from oracle import pnp_oracle
    rng = np.random.default_rng(0)
    K, X, R, t, p1, p2 = _scene(rng, 700)
    P1 = K @ np.hstack([np.eye(3), np.zeros((3, 1))])
    P2 = K @ np.hstack([R, t.reshape(3, 1)])
    x4, x3 = P.triangulate(P1, P2, p1, p2, gpu_ctx)
    w4, w3 = pnp_oracle.triangulate(P1, P2, p1, p2)
    assert np.abs(x4 - w4).max() < 1e-9 and np.abs(np.linalg.norm(x4, axis=1) - 1).max() < 1e-12 and (x4[:, 3] >= 0).all()
    assert (np.abs(x3 - w3) / np.abs(w3).max()).max() < 1e-7
    assert np.abs(x3 - X).max() < 5e-3  # float32 pixel coordinates of an exact scene: the structure comes back
    assert np.array_equal(P.triangulate_points(K, R, t, p1, p2, gpu_ctx), x3)
    # noisy correspondences: still numpy's answer
    q2 = (p2 + rng.normal(0, 0.5, p2.shape)).astype(np.float32)
    y4, y3 = P.triangulate(P1, P2, p1, q2, gpu_ctx)
    v4, v3 = pnp_oracle.triangulate(P1, P2, p1, q2)
    assert np.abs(y4 - v4).max() < 1e-9
    assert P.triangulate(P1, P2, p1[:0], p2[:0], gpu_ctx)[1].shape == (0, 3)


def test_pnp_ransac_equals_the_literal_restatement(gpu_ctx):
    import slam_cin0051_b200.pose as P
    from oracle import pnp_oracle
    rng = np.random.default_rng(1)
    K, X, R, t, p1, p2 = _scene(rng, 200)
    x = p2.astype(np.float64) + rng.normal(0, 0.3, p2.shape)
    s6 = P.pnp_sample_indices(1234, len(X), 100)
    assert s6.shape == (100, 6) and all(len(set(r)) == 6 for r in s6.tolist()) and s6.min() >= 0 and s6.max() < len(X)
    assert np.array_equal(s6, P.pnp_sample_indices(1234, len(X), 100)) and not np.array_equal(s6, P.pnp_sample_indices(1235, len(X), 100))
    for thr in (2.0, 50.0, 1e4):
        counts, Rt = P.pnp_ransac(X, x, K, s6, thr, gpu_ctx)
        wc, wRt = pnp_oracle.pnp_ransac(X, x, K, s6, thr)
        assert np.array_equal(counts, wc), (thr, np.abs(counts - wc).max())
        assert np.abs(Rt - wRt).max() < 1e-8
    # the orthogonalised block is a rotation for both signs
    Rm = Rt[:, :, :9].reshape(-1, 3, 3)
    assert np.abs(Rm @ Rm.transpose(0, 2, 1) - np.eye(3)).max() < 1e-9 and np.abs(np.linalg.det(Rm) - 1).max() < 1e-9
    # the reference's mapping defect (K never removed, column-major read-back): no consensus at the reference's own 2 px threshold
    counts, _ = P.pnp_ransac(X, x, K, s6, 2.0, gpu_ctx)
    assert counts.max() < 5
    assert P.verify_geometric_consistency(X, x, K, 100, 2.0, 5, seed=7, context=gpu_ctx) is None
    res = P.verify_geometric_consistency(X, x, K, 100, 1e4, 5, seed=7, context=gpu_ctx)
    assert res is not None and res["inliers"] >= 5 and res["relativeTransform"].shape == (4, 4)
    with pytest.raises(RuntimeError):
        P.pnp_ransac(X[:5], x[:5], K, s6[:1] % 5, 2.0, gpu_ctx)
