"""tensor-core matcher first light: knn2 vs numpy brute force on a few shapes"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import slam_cin0051_b200 as S
POP = np.array([bin(i).count("1") for i in range(256)], np.int32)

def brute(d1, d2):
    dist = POP[d1[:, None, :] ^ d2[None, :, :]].sum(-1)
    n = d2.shape[0]
    order = np.lexsort((np.broadcast_to(np.arange(n), dist.shape), dist), axis=1)[:, :2]
    return order, np.take_along_axis(dist, order, 1)

ctx = S.Context(0)
mat = S.FeatureMatcher(dict(DistanceType="HAMMING", FilterMatches=0, GoodMatchesCount=1, UseRatioTest=1, RatioTestThreshold=0.75), ctx)
bad = 0
for (n1, n2, seed) in [(128, 128, 0), (100, 300, 1), (300, 100, 2), (2000, 2000, 3), (129, 1, 4), (1, 129, 5), (1500, 2500, 6), (257, 4000, 7)]:
    rng = np.random.default_rng(seed)
    d1 = rng.integers(0, 256, (n1, 32), dtype=np.uint8)
    d2 = rng.integers(0, 256, (n2, 32), dtype=np.uint8)
    if seed % 2:  # near-duplicates and exact ties
        d2[: min(n1, n2) // 2] = d1[: min(n1, n2) // 2]
        d2[n2 // 2:] = d2[: n2 - n2 // 2]
    got = mat.knn2(d1, d2)
    order, dd = brute(d1, d2)
    ok0 = np.array_equal(got["trainIdx0"], order[:, 0]) and np.array_equal(got["distance0"], dd[:, 0])
    if n2 > 1:
        ok1 = np.array_equal(got["trainIdx1"], order[:, 1]) and np.array_equal(got["distance1"], dd[:, 1])
    else:
        ok1 = True
    print(n1, n2, "best", ok0, "second", ok1, flush=True)
    if not (ok0 and ok1):
        bad += 1
        w = np.nonzero(got["trainIdx0"] != order[:, 0])[0][:5]
        print("  first mismatches", w, got["trainIdx0"][w], order[w, 0], got["distance0"][w], dd[w, 0])
print("BAD" if bad else "ALL OK")
