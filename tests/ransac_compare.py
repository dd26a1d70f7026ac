"""How a device findEssentialMat result is compared with cv2's (shared by tests/test_gpu_scale_parity.py and
tools/soak_parity.py).  TEST INFRASTRUCTURE.

What is exact and what is not ("two-view geometry"):

* The RANSAC machinery -- cv::RNG stream, 5-subset sampling, Sampson test in double narrowed to float, strictly-greater accept
  rule, RANSACUpdateNumIters -- is reproduced EXACTLY: the device equals the oracle loop driven by the g++ build of the
  device's own minimal solver bit for bit (tests/test_gpu_essential.py, the soak), and its mask always equals the Sampson test
  re-evaluated in numpy on its own E (`self_consistent` below).
* The 5-point minimal solver is the same mathematics as OpenCV's with a different null-space basis and root finder.  On
  ordinary samples its solutions agree with cv2's to ~1e-12; on samples whose degree-10 polynomial has near-multiple real
  roots the two root finders (OpenCV: Durand-Kerner, keep |Im| <= 1e-10; here: real roots bracketed by the derivative's)
  return different SETS of solutions, and if such a sample wins, the consensus sets differ by a few points.  No independent
  fp64 implementation reproduces cv2 there (the numpy restatement does not either).
* Stated tolerance, per problem: Jaccard(mask, mask_cv2) >= 0.8 and |inliers - inliers_cv2| <= max(2, 5 % of the points);
  over a set of problems: masks IDENTICAL to cv2's on at least the stated fraction (>= 0.9 for genuine 3-D two-view geometry,
  >= 0.7 on the pure-image-translation synthetic sequences, which are a degenerate configuration for E); where the masks are
  identical, E agrees with cv2's up to sign within max(1e-9, 100 d), d = |E_numpy - E_cv2| being what the independent numpy
  restatement achieves on that problem (d / 2.2e-16 estimates the condition number of the winning sample).  E is NOT compared
  on the pure-image-translation sequences (check_E=False): the data do not determine it there -- a family of essential matrices
  explains the same flow, and identical inlier masks come with E's that are 1e-2 apart in cv2, numpy and on the device alike.
"""
import numpy as np


def e_dist(a, b):
    return float(min(np.abs(a - b).max(), np.abs(a + b).max()))


def cv2_essential(p1, p2, K4, max_iters=1000):
    import cv2
    K = np.array([[K4[0], 0, K4[2]], [0, K4[1], K4[3]], [0, 0, 1.0]])
    E, mask = cv2.findEssentialMat(p1, p2, K, cv2.RANSAC, 0.999, 1.0, max_iters)
    mask = np.zeros(len(p1), np.uint8) if mask is None else mask.ravel().astype(np.uint8)
    E = None if E is None or np.size(E) < 9 else np.asarray(E, np.float64)[:3]
    return E, mask


def compare(p1, p2, K4, dev_E, dev_mask, max_iters=1000, threshold=1.0, check_E=True):
    """Returns (ok, info): ok = the per-problem tolerance holds; info carries mask_equal / jaccard / dE for the set-level checks."""
    from oracle import essential_oracle as eo
    p1 = np.ascontiguousarray(p1, np.float32)
    p2 = np.ascontiguousarray(p2, np.float32)
    cE, cmask = cv2_essential(p1, p2, K4, max_iters)
    n = len(p1)
    good, cgood = int(dev_mask.sum()), int(cmask.sum())
    jac = float((dev_mask & cmask).sum() / max(1, (dev_mask | cmask).sum())) if (good or cgood) else 1.0
    info = {"n": n, "good": good, "cv2_good": cgood, "jaccard": jac, "mask_equal": bool(np.array_equal(dev_mask, cmask))}
    ok = jac >= 0.8 and abs(good - cgood) <= max(2, 0.05 * n)
    if good > 0:  # the device's mask is exactly the Sampson test of its own E (cv's float narrowing included)
        thr = threshold / ((K4[0] + K4[1]) / 2.0)
        err = eo.sampson_errors(np.asarray(dev_E, np.float64), eo.normalise(p1, K4), eo.normalise(p2, K4))
        info["self_consistent"] = bool(np.array_equal((err <= np.float32(thr * thr)).astype(np.uint8), dev_mask))
        ok = ok and info["self_consistent"]
    if info["mask_equal"] and cgood > 0 and cE is not None:
        d = e_dist(dev_E, cE)
        info["dE"] = d
        if d >= 1e-9 and check_E:
            nE, nmask, _ = eo.find_essential(p1, p2, K4, max_iters=max_iters)
            tol = max(1e-9, 100.0 * e_dist(nE, cE)) if nE is not None else 1.0
            info["tol"] = tol
            ok = ok and d <= tol
    return bool(ok), info


def two_view_scene(rng, n, inlier_ratio, noise_px=0.4, K4=(525.0, 525.0, 319.5, 239.5)):
    """A genuine 3-D two-view problem: points in a frustum, small rotation + translation, pixel noise, uniform outliers."""
    import cv2
    K = np.array([[K4[0], 0, K4[2]], [0, K4[1], K4[3]], [0, 0, 1.0]])
    X = np.c_[rng.uniform(-3, 3, n), rng.uniform(-2, 2, n), rng.uniform(4, 12, n)]
    R, _ = cv2.Rodrigues(rng.normal(0, 0.08, 3))
    t = rng.normal(0, 0.4, 3)
    a = (K @ X.T).T
    b = (K @ (X @ R.T + t).T).T
    a, b = a[:, :2] / a[:, 2:], b[:, :2] / b[:, 2:]
    b = b + rng.normal(0, noise_px, b.shape)
    bad = rng.random(n) > inlier_ratio
    b[bad] = rng.uniform(0, 640, (int(bad.sum()), 2))
    return a.astype(np.float32), b.astype(np.float32)
