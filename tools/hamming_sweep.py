"""BASELINE.json config 5: all-pairs Hamming matching sweep, n x n 256-bit descriptors, k = 2 (+ ratio test through
FeatureMatcher.match).  Reports G comparisons/s of the match kernel (CUDA events around the kernel, via the library's
profiler) and of the whole host call (H2D + kernel + D2H).  Run on a GPU box:  python tools/hamming_sweep.py  (--quick: one JSON line on stdout, used by bench.py's `hamming_sweep` leg)
Writes gpurun_out/hamming_sweep.json (copy under profiles/ to keep)."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import slam_cin0051_b200 as S  # noqa: E402

POP = np.array([bin(i).count("1") for i in range(256)], np.int32)


def main():
    quick = "--quick" in sys.argv
    ctx = S.Context(0)
    mat = S.FeatureMatcher(dict(DistanceType="HAMMING", FilterMatches=0, GoodMatchesCount=1, UseRatioTest=1,
                                RatioTestThreshold=0.75), ctx)
    gpopc = ctx.popc_peak()
    out = {"popc_peak_gpopc_s": gpopc, "peak_gcmp_s_4popc": gpopc / 4.0, "rows": []}
    for n in (2048, 4096, 8192, 16384, 32768, 65536):
        d1 = np.random.default_rng(0).integers(0, 256, (n, 32), dtype=np.uint8)
        d2 = np.random.default_rng(1).integers(0, 256, (n, 32), dtype=np.uint8)
        mat.knn2(d1, d2)  # warm-up (allocates the workspace)
        ctx.profile_enable(True)
        reps = 3
        t0 = time.perf_counter()
        for _ in range(reps):
            got = mat.knn2(d1, d2)
        wall = (time.perf_counter() - t0) / reps
        prof = ctx.profile_read()
        ctx.profile_enable(False)
        k_ms = prof["match"][0] / prof["match"][1]
        # spot check 48 queries against numpy
        qs = np.random.default_rng(2).choice(n, 48, replace=False)
        dist = POP[d1[qs][:, None, :] ^ d2[None, :, :]].sum(-1)
        order = np.lexsort((np.broadcast_to(np.arange(n), dist.shape), dist), axis=1)[:, :2]
        ok = bool(np.array_equal(got["trainIdx0"][qs], order[:, 0]) and np.array_equal(got["trainIdx1"][qs], order[:, 1]) and
                  np.array_equal(got["distance0"][qs], np.take_along_axis(dist, order[:, :1], 1)[:, 0]))
        gcmp = n * n / k_ms / 1e6
        row = {"n": n, "kernel_ms": k_ms, "kernel_gcmp_s": gcmp, "host_call_ms": wall * 1e3,
               "host_call_gcmp_s": n * n / wall / 1e9, "frac_of_4popc_peak": gcmp / (gpopc / 4.0), "spot_check": ok,
               "on_tensor_cores": "match_expand" in prof, "expand_ms": (prof["match_expand"][0] / reps) if "match_expand" in prof else None,
               "u8_tops": gcmp * 512 / 1e3}
        out["rows"].append(row)
        if not quick:
            print(row, flush=True)
    out["metric"] = "G Hamming cmp/s, n x n 256-bit descriptors, k=2"
    out["value_at_64k"] = out["rows"][-1]["kernel_gcmp_s"]
    if quick:  # bench.py's extra leg: one JSON line
        print(json.dumps(out), flush=True)
        return
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "hamming_sweep.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
