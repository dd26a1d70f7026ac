"""Shared inputs of tests/test_gpu_match_tc.py (imported by the test and by its child processes)."""
import numpy as np

CASES = [(128, 128, 0), (100, 300, 1), (300, 100, 2), (2000, 2000, 3), (129, 1, 4), (1, 129, 5), (1500, 2500, 6), (257, 4000, 7), (64, 65, 8),
         (2048, 2049, 9), (5000, 127, 10)]


def make(n1, n2, seed):
    rng = np.random.default_rng(seed)
    d1 = rng.integers(0, 256, (n1, 32), dtype=np.uint8)
    d2 = rng.integers(0, 256, (n2, 32), dtype=np.uint8)
    if seed % 2:  # exact duplicates (distance 0 ties) and repeated train rows (equal distances at different indices)
        h = min(n1, n2) // 2
        d2[:h] = d1[:h]
        d2[n2 // 2:] = d2[: n2 - n2 // 2]
    if seed % 3 == 0:  # extreme popcounts: all-zero and all-one descriptors (distance 256 exists)
        d1[0] = 0
        d2[-1] = 255
        d1[-1] = 255
        d2[0] = 0
    return d1, d2
