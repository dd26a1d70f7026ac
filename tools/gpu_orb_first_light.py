"""Diagnostic (not a pytest file): ORB-mode CUDA path vs the numpy ORB oracle, stage by stage."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, cv2
import slam_cin0051_b200 as s
from oracle import orb_oracle as oo
D = os.path.join(ROOT, "test", "data")
ctx = s.Context.default()
det = s.FeatureDetector(os.path.join(D, "feature_detector_orb.yml"), ctx)
for name in ["images/0000000000.png", "test_images/0.png"]:
    img = cv2.imread(os.path.join(D, name), 0)
    gk, gd = det.detect_and_compute(img)
    go = det.last_octaves(len(gk))
    levels, scales = oo.build_pyramid(img)
    quotas = oo.level_quotas(2000)
    print(name, "n gpu", len(gk))
    for l, lv in enumerate(levels):
        H, W = lv.shape
        gl = det.orb_level_image(l)
        smap = oo.fast9_scores(lv, 20)
        xs, ys, sc = oo.fast9_detect(lv, 20)
        inb = (xs >= 31) & (xs < W - 31) & (ys >= 31) & (ys < H - 31)
        xs, ys, sc = xs[inb], ys[inb], sc[inb]
        gx, gy, gv = det.orb_stage(0, l)
        same = len(gx) == len(xs) and np.array_equal(gx, xs) and np.array_equal(gy, ys) and np.array_equal(gv.astype(np.int32), sc)
        print(f"  L{l} {W}x{H} level image eq {np.array_equal(gl, lv)} cand oracle {len(xs)} gpu {len(gx)} identical {same}")
        if not same:
            oset = set(zip(xs.tolist(), ys.tolist())); gset = set(zip(gx.tolist(), gy.tolist()))
            extra = sorted(gset - oset, key=lambda p: (p[1], p[0])); missing = sorted(oset - gset, key=lambda p: (p[1], p[0]))
            print("    extra", len(extra), "missing", len(missing))
            sval = {(x, y): v for x, y, v in zip(gx.tolist(), gy.tolist(), gv.tolist())}
            for (x, y) in extra[:6]:
                print("    extra", (x, y), "gpu score", sval[(x, y)], "oracle 3x3\n", smap[y - 1:y + 2, x - 1:x + 2])
            ex = np.array(extra)
            if len(ex):
                print("    extra x%32 hist", np.bincount(ex[:, 0] % 32, minlength=32).tolist())
                print("    extra y%32 hist", np.bincount(ex[:, 1] % 32, minlength=32).tolist())
    t = time.time(); det.detect_and_compute(img); print(f"  gpu orb {1e3*(time.time()-t):.2f} ms")
