"""Recovers OpenCV's rBRIEF sampling pattern (bit_pattern_31_, 256 x 4 int32) from the cv2 wheel's binary.

The pattern is OpenCV data (features2d/src/orb.cpp), not part of the reference repository; OpenCV is an
un-vendored dependency of the reference (conanfile.txt: opencv/4.12.0).  The table is located by its first
eight values.  Output: slam_cin0051_b200/orb_bit_pattern_31.npy (committed; 4 KB).
"""
import os
import sys

import cv2
import numpy as np

needle = np.array([8, -3, 9, 5, 4, 2, 7, -12], np.int32).tobytes()
so = [os.path.join(os.path.dirname(cv2.__file__), f) for f in os.listdir(os.path.dirname(cv2.__file__)) if f.endswith(".so")]
for path in so:
    blob = open(path, "rb").read()
    at = blob.find(needle)
    if at >= 0:
        pat = np.frombuffer(blob[at:at + 256 * 4 * 4], np.int32).reshape(256, 4).copy()
        assert np.abs(pat).max() <= 15 and blob.find(needle, at + 1) < 0
        out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "slam_cin0051_b200", "orb_bit_pattern_31.npy")
        np.save(out, pat)
        print("found in", path, "at", at, "->", out, pat[:2].tolist(), pat[-1].tolist())
        sys.exit(0)
print("pattern not found")
sys.exit(1)
