"""Generates tests/golden/orb_*.npz: outputs of OpenCV itself (cv2.ORB_create / BFMatcher.knnMatch /
findEssentialMat) on the reference's fixtures and on the synthetic KITTI-shape frames.

OpenCV is the un-vendored dependency in which the "ORB mode" arithmetic lives (reference conanfile.txt:2 pins
opencv/4.12.0; this image has the cv2 4.13.0 wheel).  The vectors pin both the numpy oracle
(oracle/orb_oracle.py) and the CUDA path.  Keypoints are stored in canonical order (octave, y, x) because
cv2's in-level order comes from nth_element/partition and is implementation defined; knn indices refer to
that canonical order.  Run from the repo root:  python tools/make_golden_orb.py
"""
import os
import sys

import cv2
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from slam_cin0051_b200.synth import make_sequence  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
DATA = os.path.join(ROOT, "test", "data")


def orb_canonical(img, nfeatures=2000, nlevels=8, scale=1.2, fast=20):
    orb = cv2.ORB_create(nfeatures=nfeatures, scaleFactor=scale, nlevels=nlevels, edgeThreshold=31, firstLevel=0,
                         WTA_K=2, scoreType=cv2.ORB_HARRIS_SCORE, patchSize=31, fastThreshold=fast)
    kps, desc = orb.detectAndCompute(img, None)
    if desc is None:
        desc = np.zeros((0, 32), np.uint8)
    key = np.array([(k.octave, k.pt[1], k.pt[0]) for k in kps], np.float64).reshape(-1, 3)
    order = np.lexsort((key[:, 2], key[:, 1], key[:, 0]))
    f = lambda g: np.array([g(kps[i]) for i in order], np.float32)
    return {"x": f(lambda k: k.pt[0]), "y": f(lambda k: k.pt[1]), "size": f(lambda k: k.size),
            "angle": f(lambda k: k.angle), "response": f(lambda k: k.response),
            "octave": np.array([kps[i].octave for i in order], np.int8), "desc": desc[order]}


def knn2(d1, d2):
    m = cv2.BFMatcher(cv2.NORM_HAMMING).knnMatch(d1, d2, k=2)
    idx = np.array([[a.trainIdx, b.trainIdx] for a, b in m], np.int32)
    dist = np.array([[a.distance, b.distance] for a, b in m], np.float32)
    return idx, dist


def main():
    cv2.setNumThreads(1)
    os.makedirs(OUT, exist_ok=True)
    syn = make_sequence(376, 1241, 2, pitch_px=14, seed=0)
    images = {
        "kitti0": cv2.imread(os.path.join(DATA, "images/0000000000.png"), 0),
        "kitti1": cv2.imread(os.path.join(DATA, "images/0000000001.png"), 0),
        "tum0": cv2.imread(os.path.join(DATA, "test_images/0.png"), 0),
        "synK0": syn[0], "synK1": syn[1],
    }
    res = {}
    for name, img in images.items():
        res[name] = orb_canonical(img)
        np.savez_compressed(os.path.join(OUT, f"orb_{name}.npz"), cv2_version=cv2.__version__, **res[name])
        print(name, img.shape, len(res[name]["x"]))
    for a, b in (("kitti0", "kitti1"), ("synK0", "synK1")):
        idx, dist = knn2(res[a]["desc"], res[b]["desc"])
        np.savez_compressed(os.path.join(OUT, f"knn2_{a}_{b}.npz"), idx=idx, dist=dist, cv2_version=cv2.__version__)
        print("knn2", a, b, idx.shape)
    # small-parameter case: 4 levels, scale 1.5, 300 features, FAST threshold 30 (TUM frame)
    small = orb_canonical(images["tum0"], nfeatures=300, nlevels=4, scale=1.5, fast=30)
    np.savez_compressed(os.path.join(OUT, "orb_tum0_n300_l4_s15_t30.npz"), cv2_version=cv2.__version__, **small)
    print("small", len(small["x"]))


if __name__ == "__main__":
    main()
