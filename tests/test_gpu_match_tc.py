"""The tensor-core Hamming matcher (csrc/match_tc.cu) against the integer-pipe kernels of match.cu and numpy: the same best / second
neighbour (index and distance) for every query -- ties, duplicates, ragged sizes, one-row sets, sliced train sets -- and the check that
the tensor-core path is the one that ran (its operand-widening kernel shows up in the library's profile).
Reference semantics: feature_matcher.cpp:132-141 (strict <, lowest index wins), :153-173."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
POP = np.array([bin(i).count("1") for i in range(256)], np.int32)

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from match_cases import CASES, make  # noqa: E402


def brute(d1, d2):
    dist = POP[d1[:, None, :] ^ d2[None, :, :]].sum(-1)
    n = d2.shape[0]
    order = np.lexsort((np.broadcast_to(np.arange(n), dist.shape), dist), axis=1)[:, :2]
    return order, np.take_along_axis(dist, order, 1)


CHILD = r"""
import json, sys, numpy as np
sys.path.insert(0, %r)
sys.path.insert(0, %r)
import slam_cin0051_b200 as S
from match_cases import CASES, make
ctx = S.Context(0)
mat = S.FeatureMatcher(dict(DistanceType="HAMMING", FilterMatches=0, GoodMatchesCount=1, UseRatioTest=1, RatioTestThreshold=0.75), ctx)
out = {}
ctx.profile_enable(True)
for (n1, n2, seed) in CASES:
    d1, d2 = make(n1, n2, seed)
    g = mat.knn2(d1, d2)
    out[str(seed)] = {k: np.asarray(g[k]).tolist() for k in ("trainIdx0", "trainIdx1", "distance0", "distance1")}
out["kernels"] = sorted(ctx.profile_read().keys())
json.dump(out, open(sys.argv[1], "w"))
"""


def run_child(tmp_path, tc):
    path = tmp_path / f"knn_{tc}.json"
    env = dict(os.environ, SLAMCU_MATCH_TC=tc, PYTHONPATH=ROOT)
    subprocess.run([sys.executable, "-c", CHILD % (ROOT, os.path.join(ROOT, "tests")), str(path)], check=True, env=env, cwd=ROOT, timeout=300)
    return json.load(open(path))


def test_tensor_core_path_equals_integer_pipe_and_numpy(tmp_path):
    tcs = run_child(tmp_path, "1")
    ints = run_child(tmp_path, "0")
    assert "match_expand" in tcs["kernels"], "the tensor-core path did not run"
    assert "match_expand" not in ints["kernels"]
    for (n1, n2, seed) in CASES:
        a, b = tcs[str(seed)], ints[str(seed)]
        for k in ("trainIdx0", "trainIdx1", "distance0", "distance1"):
            assert a[k] == b[k], (n1, n2, seed, k)
        d1, d2 = make(n1, n2, seed)
        order, dd = brute(d1, d2)
        assert np.array_equal(np.asarray(a["trainIdx0"]), order[:, 0]) and np.array_equal(np.asarray(a["distance0"]), dd[:, 0]), (n1, n2)
        if n2 > 1:
            assert np.array_equal(np.asarray(a["trainIdx1"]), order[:, 1]) and np.array_equal(np.asarray(a["distance1"]), dd[:, 1]), (n1, n2)


def test_sliced_train_sets_agree():
    import slam_cin0051_b200 as S
    ctx = S.Context(0)
    mat = S.FeatureMatcher(dict(DistanceType="HAMMING", FilterMatches=0, GoodMatchesCount=1, UseRatioTest=1, RatioTestThreshold=0.75), ctx)
    d1, d2 = make(700, 5000, 11)
    order, dd = brute(d1, d2)
    for slices in (1, 3, 7, 32):
        mat.set_train_slices(slices)
        g = mat.knn2(d1, d2)
        assert np.array_equal(g["trainIdx0"], order[:, 0]) and np.array_equal(g["trainIdx1"], order[:, 1]), slices
        assert np.array_equal(g["distance0"], dd[:, 0]) and np.array_equal(g["distance1"], dd[:, 1]), slices
