"""CPU checks of slam_cin0051_b200/csrc/fivept.cuh (the five-point solver the RANSAC kernel runs, compiled for the
host by tests/native/host_exact.cpp): its real-root finder against numpy.roots, its model sets against the oracle's
solver, and cv::findEssentialMat's RANSAC loop driven by it against the committed cv2 outputs."""
import ctypes as C
import glob
import os
import subprocess

import numpy as np
import pytest

from oracle import essential_oracle as eo

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = sorted(glob.glob(os.path.join(HERE, "golden", "essential_*.npz")))


@pytest.fixture(scope="module")
def hx():
    so = os.path.join(HERE, "native", "libhost_exact.so")
    src = os.path.join(HERE, "native", "host_exact.cpp")
    hdrs = [os.path.join(HERE, "..", "slam_cin0051_b200", "csrc", h) for h in ("exact.cuh", "fivept.cuh")]
    if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(p) for p in [src, *hdrs]):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-o", so, src])
    return C.CDLL(so)


def k4(K):
    return (K[0, 0], K[1, 1], K[0, 2], K[1, 2])


def e_diff(a, b):
    return min(np.abs(a - b).max(), np.abs(a + b).max())


def host_five_point(hx, x1, x2):
    a = np.ascontiguousarray(x1, np.float64).reshape(1, 5, 2)
    b = np.ascontiguousarray(x2, np.float64).reshape(1, 5, 2)
    models = np.zeros((1, 10, 9))
    counts = np.zeros(1, np.int32)
    hx.hx_five_point(a.ctypes.data, b.ctypes.data, 1, models.ctypes.data, counts.ctypes.data)
    return [models[0, k].reshape(3, 3).copy() for k in range(counts[0])]


def test_real_roots_equal_numpy_roots(hx):
    rng = np.random.default_rng(0)
    for trial in range(2000):
        kind = trial % 5
        if kind == 0:
            c = rng.normal(size=11)
        elif kind == 1:  # coefficients over many orders of magnitude
            c = rng.normal(size=11) * 10.0 ** rng.integers(-4, 5, 11)
        elif kind == 2:  # prescribed real roots (small and large) + complex pairs
            r = list(rng.normal(size=4) * 10.0 ** rng.integers(-3, 4, 4))
            z = rng.normal(size=3) + 1j * rng.normal(size=3)
            c = np.real(np.poly(r + list(z) + list(np.conj(z)))) * rng.normal()
        elif kind == 3:  # leading zeros: lower degree
            c = np.concatenate([[0.0, 0.0], rng.normal(size=9)])
        else:  # low degrees
            c = rng.normal(size=rng.integers(2, 6))
        c = np.ascontiguousarray(c, np.float64)
        out = np.zeros(10)
        n = hx.hx_real_roots(c.ctypes.data, len(c), out.ctypes.data)
        mine = np.sort(out[:n])
        ref = np.roots(np.trim_zeros(c, "f"))
        want = np.sort(ref[np.abs(ref.imag) <= 1e-10 * np.maximum(1, np.abs(ref.real))].real)
        assert len(mine) == len(want), (trial, mine, ref)
        if len(mine):
            assert np.max(np.abs(mine - want) / np.maximum(1, np.abs(want))) < 1e-7, (trial, mine, want)


def test_real_roots_special_cases(hx):
    def roots(c):
        c = np.ascontiguousarray(c, np.float64)
        out = np.zeros(10)
        return sorted(out[: hx.hx_real_roots(c.ctypes.data, len(c), out.ctypes.data)])
    assert roots([1.0, 0.0, -1.0]) == [-1.0, 1.0]                       # roots on the unit interval's ends
    assert roots([1.0, -3.0, 2.0, 0.0]) == [0.0, 1.0, 2.0]              # a root at zero (reversed polynomial loses a degree)
    assert roots([2.0, 0.0, 2.0]) == []                                 # no real root
    assert roots([0.0, 0.0, 5.0]) == [] and roots([0.0, 2.0, -8.0]) == [4.0]
    got = roots(np.poly([1e-6, 0.5, 3.0, 1e5]))
    assert np.allclose(got, [1e-6, 0.5, 3.0, 1e5], rtol=1e-9, atol=0)


def test_five_point_models_equal_oracle_solver(hx):
    total = close = 0
    for gi, path in enumerate(GOLD):
        g = np.load(path)
        x1, x2 = eo.normalise(g["p1"], k4(g["K"])), eo.normalise(g["p2"], k4(g["K"]))
        rng = np.random.default_rng(gi)
        for _ in range(25):
            idx = rng.choice(len(x1), 5, replace=False)
            got = host_five_point(hx, x1[idx], x2[idx])
            want = eo.five_point(x1[idx], x2[idx])
            assert abs(len(got) - len(want)) <= 1
            for M in got:
                total += 1
                close += min((e_diff(M, W) for W in want), default=9.0) < 1e-8
                assert abs(np.sqrt((M * M).sum()) - 1.0) < 1e-12
                assert max(abs(np.array([*x2[j], 1.0]) @ M @ np.array([*x1[j], 1.0])) for j in idx) < 1e-12
    # the remainder are ill-conditioned samples (nearly double roots) where both solvers are off by the same order
    assert total > 500 and close >= 0.97 * total


@pytest.mark.parametrize("path", GOLD, ids=[os.path.basename(p)[10:-4] for p in GOLD])
def test_ransac_with_the_kernel_solver_equals_cv2_golden(hx, path):
    g = np.load(path)
    E, mask, good = eo.find_essential(g["p1"], g["p2"], k4(g["K"]), solver=lambda a, b: host_five_point(hx, a, b))
    assert good == int(g["mask"].sum()) and np.array_equal(mask, g["mask"])
    assert e_diff(E, g["E"]) < 1e-9
