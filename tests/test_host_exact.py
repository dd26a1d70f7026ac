"""CPU checks of slam_cin0051_b200/csrc/exact.cuh (compiled for the host by tests/native/host_exact.cpp):
the glibc float-libm ports and the libstdc++ sort-permutation emulations that the CUDA kernels use must be
bit-identical to the real libm / libstdc++ of this image (the libraries the reference would be built against)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def hx():
    so = os.path.join(HERE, "native", "libhost_exact.so")
    src = os.path.join(HERE, "native", "host_exact.cpp")
    hdrs = [os.path.join(HERE, "..", "slam_cin0051_b200", "csrc", h) for h in ("exact.cuh", "fivept.cuh")]
    if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(p) for p in [src, *hdrs]):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-o", so, src])
    L = C.CDLL(so)
    L.hx_check_atanf_bits.restype = C.c_longlong
    L.hx_check_atanf_bits.argtypes = [C.c_uint32, C.c_longlong]
    L.hx_check_sincosf_bits.restype = C.c_longlong
    L.hx_check_sincosf_bits.argtypes = [C.c_uint32, C.c_longlong]
    L.hx_check_atan2f.restype = C.c_longlong
    L.hx_check_atan2f.argtypes = [C.c_void_p, C.c_void_p, C.c_longlong]
    return L


def test_atanf_port_bitexact_samples(hx):
    # every branch of the fdlibm code: tiny, <0.4375, the four reduction intervals, huge; both signs
    for start in (0x00000001, 0x30800000, 0x3e000000, 0x3ee00000, 0x3f300000, 0x3f980000, 0x401c0000, 0x4b800000, 0x7f000000):
        for sign in (0, 0x80000000):
            assert hx.hx_check_atanf_bits(start | sign, 2_000_000) == 0


def test_atan2f_port_on_integer_moments(hx):
    rng = np.random.default_rng(0)
    n = 3_000_000
    y = rng.integers(-2_700_000, 2_700_000, n).astype(np.float32)
    x = rng.integers(-2_700_000, 2_700_000, n).astype(np.float32)
    y[:1000] = 0
    x[1000:2000] = 0
    x[2000:3000] = 1
    y[3000:4000] = -0.0
    x[3000:3500] = -5
    assert hx.hx_check_atan2f(y.ctypes.data, x.ctypes.data, n) == 0
    y2 = rng.integers(-30, 30, 200_000).astype(np.float32)
    x2 = rng.integers(-30, 30, 200_000).astype(np.float32)
    assert hx.hx_check_atan2f(y2.ctypes.data, x2.ctypes.data, len(y2)) == 0


def test_sincosf_port_bitexact_over_angle_range(hx):
    # |a| <= pi is all the reference feeds (angle * DEG2RAD); cover [2^-13, 3.2] densely in both signs
    for start in (0x39000000, 0x3c000000, 0x3f000000, 0x3f490fdb - 1000, 0x3fc00000, 0x40400000):
        for sign in (0, 0x80000000):
            assert hx.hx_check_sincosf_bits(start | sign, 3_000_000) == 0
    assert hx.hx_check_sincosf_bits(0x40490fdb - 500_000, 1_000_000) == 0  # around pi


def _keys(scores, shift):
    n = len(scores)
    return ((scores.astype(np.uint32) << shift) | np.arange(n, dtype=np.uint32)).astype(np.uint32)


def _median3_killer(n):
    # Musser's median-of-3 killer: drives introsort into its depth-limit heap-sort fallback
    k = n // 2
    a = np.zeros(n, np.int64)
    for i in range(k):
        a[i] = i + 1 if i % 2 == 0 else k + i + 1
        a[k + i] = 2 * (i + 1)
    return a


@pytest.mark.parametrize("n", [0, 1, 2, 15, 16, 17, 18, 33, 100, 1000, 11329, 40000])
def test_std_sort_emulations_match_libstdcxx(hx, n):
    rng = np.random.default_rng(n)
    inputs = [rng.integers(0, hi, n) for hi in (4081, 50, 3)]
    inputs += [np.arange(n)[::-1] % 4081, np.arange(n) % 4081, np.zeros(n, np.int64)]
    if n >= 100:
        inputs.append(_median3_killer(n) % 4081)
        inputs.append(4080 - (_median3_killer(n) % 4081))
    for sc in inputs:
        k = _keys(np.asarray(sc), 20)
        ref = k.copy()
        hx.hx_ref_sort_keys_desc(ref.ctypes.data_as(C.c_void_p), n)
        ser = k.copy()
        hx.hx_sort_keys_desc(ser.ctypes.data_as(C.c_void_p), n)  # serial emulation (exact.cuh std_sort)
        par = k.copy()
        hx.hx_model_sort_keys_desc(par.ctypes.data_as(C.c_void_p), n)  # scalar model of the parallel CUDA formulation
        assert np.array_equal(ser, ref)
        assert np.array_equal(par, ref)


@pytest.mark.parametrize("n", [1, 2, 19, 20, 21, 64, 500, 5000])
def test_std_partial_sort_emulation(hx, n):
    rng = np.random.default_rng(n + 7)
    for hi in (257, 8, 2):
        d = rng.integers(0, hi, n)
        k = _keys(d, 16)
        for mid in (1, 5, 20, n):
            if mid > n:
                continue
            a, b = k.copy(), k.copy()
            hx.hx_partial_sort_asc(a.ctypes.data_as(C.c_void_p), mid, n)
            hx.hx_ref_partial_sort_asc(b.ctypes.data_as(C.c_void_p), mid, n)
            assert np.array_equal(a[:mid], b[:mid])
        a, b = k.copy(), k.copy()
        hx.hx_sort_asc(a.ctypes.data_as(C.c_void_p), n)
        hx.hx_ref_sort_asc(b.ctypes.data_as(C.c_void_p), n)
        assert np.array_equal(a, b)
