"""Generates tests/golden/essential_*.npz: cv2.findEssentialMat(p1, p2, K, RANSAC, 0.999, 1.0) outputs (E, mask) for
synthetic two-view scenes and for ORB matches of the reference's fixture pairs.  OpenCV is the un-vendored dependency
the reference calls at src/frontend/pose_estimator.cpp:42.  Run from the repo root: python tools/make_golden_essential.py
"""
import os
import sys

import cv2
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
OUT = os.path.join(ROOT, "tests", "golden")
DATA = os.path.join(ROOT, "test", "data")


def scene(n, noise, outliers, seed, K):
    r = np.random.default_rng(seed)
    X = np.stack([r.uniform(-4, 4, n), r.uniform(-3, 3, n), r.uniform(4, 12, n)], 1)
    R = cv2.Rodrigues(np.array([0.02, 0.05, -0.01]) * (1 + seed % 3))[0]
    t = np.array([0.3, 0.02, 0.05])
    p1 = (K @ X.T).T
    p1 = p1[:, :2] / p1[:, 2:]
    X2 = (R @ X.T).T + t
    p2 = (K @ X2.T).T
    p2 = p2[:, :2] / p2[:, 2:]
    p1 += r.normal(0, noise, p1.shape)
    p2 += r.normal(0, noise, p2.shape)
    k = int(outliers * n)
    p2[:k] = r.uniform(0, 640, (k, 2))
    return p1.astype(np.float32), p2.astype(np.float32)


def orb_matches(a, b):
    orb = cv2.ORB_create(nfeatures=2000)
    ka, da = orb.detectAndCompute(a, None)
    kb, db = orb.detectAndCompute(b, None)
    m = cv2.BFMatcher(cv2.NORM_HAMMING).knnMatch(da, db, k=2)
    good = [x for x, y in m if x.distance < 0.75 * y.distance]
    p1 = np.array([ka[g.queryIdx].pt for g in good], np.float32)
    p2 = np.array([kb[g.trainIdx].pt for g in good], np.float32)
    return p1, p2


def main():
    cv2.setNumThreads(1)
    os.makedirs(OUT, exist_ok=True)
    Kt = np.array([[525.0, 0, 319.5], [0, 525.0, 239.5], [0, 0, 1]])
    Kk = np.array([[984.2439, 0, 690.0], [0, 980.8141, 233.1966], [0, 0, 1]])  # test/data/camera.yml K0
    cases = {
        "syn300": (*scene(300, 0.5, 0.2, 0, Kt), Kt),
        "syn1000": (*scene(1000, 0.3, 0.4, 1, Kt), Kt),
        "syn60": (*scene(60, 1.0, 0.5, 2, Kt), Kt),
        "syn2000_clean": (*scene(2000, 0.1, 0.05, 3, Kt), Kt),
        "tum01": (*orb_matches(cv2.imread(os.path.join(DATA, "test_images/0.png"), 0),
                               cv2.imread(os.path.join(DATA, "test_images/1.png"), 0)), Kt),
        "kitti01": (*orb_matches(cv2.imread(os.path.join(DATA, "images/0000000000.png"), 0),
                                 cv2.imread(os.path.join(DATA, "images/0000000001.png"), 0)), Kk),
    }
    for name, (p1, p2, K) in cases.items():
        E, mask = cv2.findEssentialMat(p1, p2, K, method=cv2.RANSAC, prob=0.999, threshold=1.0)
        # R, t: the reference's simpleRecoverPose restated over cv2.SVDecomp (= cv::SVD), on ALL correspondences
        from oracle import essential_oracle as eo
        R, t, front = eo.simple_recover_pose(E, p1, p2, (K[0, 0], K[1, 1], K[0, 2], K[1, 2]), backend=cv2)
        np.savez_compressed(os.path.join(OUT, f"essential_{name}.npz"), p1=p1, p2=p2, K=K, E=E, mask=mask.ravel().astype(np.uint8),
                            R=R, t=t, front=np.array(front, np.int32), cv2_version=cv2.__version__)
        print(name, len(p1), "inliers", int(mask.sum()), E.shape)


if __name__ == "__main__":
    main()
