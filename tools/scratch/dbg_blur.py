import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import slam_cin0051_b200 as S
from slam_cin0051_b200.synth import make_sequence
ctx = S.Context(0)
det = S.FeatureDetector(os.path.join(ROOT, "test/data/feature_detector_orb.yml"), ctx)
for shape in ((240, 320), (376, 1241), (512, 1392)):
    img = make_sequence(shape[0], shape[1], 1, 14, seed=1)[0]
    try:
        k, d = det.detect_and_compute(img)
        print(shape, len(k), "ok", flush=True)
    except Exception as e:
        print(shape, "FAIL", e, flush=True)
        break
