"""CPU model of the tensor-core matcher's arithmetic (slam_cin0051_b200/csrc/match_tc.cu), in numpy with the kernel's own 32-bit
operations: the nibble -> four-bytes multiply of expand_bits_kernel, hamming = popc(q) + popc(t) - 2 <q, t> on the widened rows,
the packed key constants and the one-multiply-add key of the epilogue (mod 2^32), and the 2.5-operation top-2 merge.  The result
must be the (distance, lowest index first) top-2 of a brute force -- the strict-< rule of updateBestMatches
(feature_matcher.cpp:132-141).  The GPU tests check the kernel against the same brute force; this one pins the formulas."""
import numpy as np

POP = np.array([bin(i).count("1") for i in range(256)], np.int64)
U32 = np.uint64(0xFFFFFFFF)


def widen(desc_words):
    """uint32 [n][8] -> uint8 [n][256], bit b of a nibble -> byte b of a word: ((nibble * 0x00204081) & 0x01010101)"""
    n = desc_words.shape[0]
    out = np.zeros((n, 64), np.uint32)
    for w in range(8):
        for nib in range(8):
            nibble = (desc_words[:, w] >> np.uint32(4 * nib)) & np.uint32(0xF)
            prod = (nibble.astype(np.uint64) * np.uint64(0x00204081)) & U32  # the 16 partial products land on distinct bits: no carries
            out[:, w * 8 + nib] = (prod & np.uint64(0x01010101)).astype(np.uint32)
    return out.view(np.uint8).reshape(n, 256)


def test_widening_is_one_byte_per_bit():
    rng = np.random.default_rng(0)
    d = rng.integers(0, 1 << 32, (50, 8), dtype=np.uint64).astype(np.uint32)
    x = widen(d)
    assert set(np.unique(x)) <= {0, 1}
    assert np.array_equal(x.sum(1), POP[d.view(np.uint8)].sum(1))
    # byte k of the row is bit k of the descriptor (little-endian words)
    bits = np.unpackbits(d.view(np.uint8), axis=1, bitorder="little")
    assert np.array_equal(x, bits)


def tc_top2(dq, dt):
    xq, xt = widen(dq).astype(np.int64), widen(dt).astype(np.int64)
    dot = xq @ xt.T  # the u8 x u8 -> s32 GEMM
    pq, pt = xq.sum(1), xt.sum(1)
    nt = dt.shape[0]
    ck = (((pt + 512) << 20) | np.arange(nt)).astype(np.uint64)  # key constant of a train row
    mul = np.uint64((-(2 << 20)) & 0xFFFFFFFF)
    keys = (dot.astype(np.uint64) * mul + ck[None, :]) & U32  # ONE mad.lo.u32 per table entry
    best = np.full(dq.shape[0], 0xFFFFFFFF, np.uint64)
    second = best.copy()
    for j in range(0, nt - 1, 2):  # pairs ordered first, then a 3-input merge
        lo, hi = np.minimum(keys[:, j], keys[:, j + 1]), np.maximum(keys[:, j], keys[:, j + 1])
        second = np.minimum(np.minimum(second, np.maximum(best, lo)), hi)
        best = np.minimum(best, lo)
    if nt % 2:
        k = keys[:, nt - 1]
        second = np.minimum(second, np.maximum(best, k))
        best = np.minimum(best, k)
    dist = lambda key: (key >> np.uint64(20)).astype(np.int64) - 512 + pq
    idx = lambda key: (key & np.uint64(0xFFFFF)).astype(np.int64)
    return idx(best), dist(best), idx(second), dist(second)


def test_key_arithmetic_gives_the_reference_top2():
    rng = np.random.default_rng(1)
    for nq, nt in ((40, 1), (33, 2), (64, 129), (50, 300)):
        dq = rng.integers(0, 1 << 32, (nq, 8), dtype=np.uint64).astype(np.uint32)
        dt = rng.integers(0, 1 << 32, (nt, 8), dtype=np.uint64).astype(np.uint32)
        if nt > 4:
            dt[: min(nq, nt) // 2] = dq[: min(nq, nt) // 2]  # distance 0
            dt[nt // 2:] = dt[: nt - nt // 2]                  # equal distances at different indices: the lowest index must win
            dq[0] = 0
            dt[-1] = 0xFFFFFFFF                               # distance 256 exists
        dist = POP[dq.view(np.uint8)[:, None, :] ^ dt.view(np.uint8)[None, :, :]].sum(-1)
        order = np.lexsort((np.broadcast_to(np.arange(nt), dist.shape), dist), axis=1)
        b_idx, b_d, s_idx, s_d = tc_top2(dq, dt)
        assert np.array_equal(b_idx, order[:, 0]) and np.array_equal(b_d, np.take_along_axis(dist, order[:, :1], 1)[:, 0])
        if nt > 1:
            assert np.array_equal(s_idx, order[:, 1]) and np.array_equal(s_d, np.take_along_axis(dist, order[:, 1:2], 1)[:, 0])
