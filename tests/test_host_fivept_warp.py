"""The WARP-COOPERATIVE five-point solver that the RANSAC kernels run (slam_cin0051_b200/csrc/fivept_warp.cuh), executed
on the CPU: tests/native/warp_solver_host.cpp compiles the device source against tests/native/simt_emu.hpp (one warp =
32 threads, every *_sync intrinsic = barrier + exchange).  Checked against the g++ build of the thread-per-sample solver
(fivept.cuh) and, through the oracle's RANSAC loop, against the committed cv2 outputs."""
import ctypes as C
import glob
import os
import subprocess

import numpy as np
import pytest

from oracle import essential_oracle as eo

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = sorted(glob.glob(os.path.join(HERE, "golden", "essential_*.npz")))
CSRC = os.path.join(HERE, "..", "slam_cin0051_b200", "csrc")


def _build(so, src, extra, deps):
    if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(p) for p in [src, *deps]):
        subprocess.check_call(["g++", "-O2", *extra, "-fPIC", "-shared", "-Wno-unknown-pragmas", "-o", so, src])
    return C.CDLL(so)


@pytest.fixture(scope="module")
def libs():
    n = os.path.join(HERE, "native")
    hdrs = [os.path.join(CSRC, h) for h in ("exact.cuh", "fivept.cuh", "fivept_warp.cuh")]
    hx = _build(os.path.join(n, "libhost_exact.so"), os.path.join(n, "host_exact.cpp"), ["-std=c++17"], hdrs[:2])
    hw = _build(os.path.join(n, "libwarp_solver_host.so"), os.path.join(n, "warp_solver_host.cpp"), ["-std=c++20", "-pthread"],
                hdrs + [os.path.join(n, "simt_emu.hpp")])
    return hx, hw


def k4(K):
    return (K[0, 0], K[1, 1], K[0, 2], K[1, 2])


def e_diff(a, b):
    return min(np.abs(a - b).max(), np.abs(a + b).max())


def solve(lib, fn, x1, x2):
    a = np.ascontiguousarray(x1, np.float64).reshape(-1, 5, 2)
    b = np.ascontiguousarray(x2, np.float64).reshape(-1, 5, 2)
    models = np.zeros((len(a), 10, 9))
    counts = np.zeros(len(a), np.int32)
    getattr(lib, fn)(a.ctypes.data, b.ctypes.data, len(a), models.ctypes.data, counts.ctypes.data)
    return [[models[i, k].reshape(3, 3).copy() for k in range(counts[i])] for i in range(len(a))]


def test_warp_solver_equals_thread_solver(libs):
    hx, hw = libs
    diffs, same, samples = [], 0, 0
    for gi, path in enumerate(GOLD):
        g = np.load(path)
        x1, x2 = eo.normalise(g["p1"], k4(g["K"])), eo.normalise(g["p2"], k4(g["K"]))
        rng = np.random.default_rng(40 + gi)
        idx = np.stack([rng.choice(len(x1), 5, replace=False) for _ in range(10)])
        want = solve(hx, "hx_five_point", x1[idx], x2[idx])
        got = solve(hw, "hw_five_point_warp", x1[idx], x2[idx])
        for i in range(len(idx)):
            samples += 1
            same += len(got[i]) == len(want[i])
            for M in got[i]:
                diffs.append(min((e_diff(M, W) for W in want[i]), default=9.0))
                assert abs(np.sqrt((M * M).sum()) - 1.0) < 1e-12
                assert max(abs(np.array([*x2[j], 1.0]) @ M @ np.array([*x1[j], 1.0])) for j in idx[i]) < 1e-11
    diffs = np.array(diffs)
    # identical mathematics and basis order; a few ill-conditioned samples (nearly double roots) differ in the digits
    # both solvers lose
    assert len(diffs) > 200 and same >= 0.97 * samples
    assert np.median(diffs) < 1e-11 and (diffs < 1e-8).mean() >= 0.95


@pytest.mark.parametrize("name", ["kitti01", "syn300", "tum01"])
def test_ransac_with_the_warp_solver_equals_cv2_golden(libs, name):
    _, hw = libs
    g = np.load([p for p in GOLD if p.endswith(f"essential_{name}.npz")][0])
    E, mask, good = eo.find_essential(g["p1"], g["p2"], k4(g["K"]), solver=lambda a, b: solve(hw, "hw_five_point_warp", a, b)[0])
    assert good == int(g["mask"].sum()) and np.array_equal(mask, g["mask"])
    assert e_diff(E, g["E"]) < 1e-9


def test_warp_solver_on_degenerate_samples(libs):
    """Inputs the RANSAC kernels do meet (scene cuts, pure translation over a plane, repeated points): the device routine
    must terminate and return at most ten finite unit-norm matrices that satisfy the five epipolar constraints (how many
    real roots a degenerate polynomial has is decided by rounding, so the count is not compared with the thread solver)."""
    _, hw = libs
    rng = np.random.default_rng(7)
    base = rng.uniform(-0.5, 0.5, (5, 2))
    cases = {
        "identical points": (np.repeat(base[:1], 5, 0), np.repeat(base[:1], 5, 0)),
        "zero motion": (base, base.copy()),
        "pure translation of a plane": (base, base + np.array([0.01, 0.0])),
        "two coincident correspondences": (np.vstack([base[:4], base[:1]]), np.vstack([base[:4] + 0.02, base[:1] + 0.02])),
        "collinear": (np.stack([np.linspace(-0.4, 0.4, 5), np.zeros(5)], 1), np.stack([np.linspace(-0.38, 0.41, 5), np.full(5, 0.01)], 1)),
        "huge coordinates": (base * 1e6, base * 1e6 + 3.0),
        "tiny coordinates": (base * 1e-9, base * 1e-9 + 1e-12),
        "integer pixels, unnormalised": (np.rint(base * 600 + 320), np.rint(base * 600 + 322)),
    }
    for name, (a, b) in cases.items():
        got = solve(hw, "hw_five_point_warp", a, b)[0]
        assert len(got) <= 10, name
        scale = max(1.0, float(np.abs(a).max())) * max(1.0, float(np.abs(b).max()))
        for M in got:
            assert np.isfinite(M).all() and abs(np.sqrt((M * M).sum()) - 1.0) < 1e-9, name
            assert max(abs(np.array([*b[j], 1.0]) @ M @ np.array([*a[j], 1.0])) for j in range(5)) < 1e-9 * scale, name
