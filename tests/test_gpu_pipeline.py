"""The pipelined host-buffer entry point (slamcu_sequence_process: chunked H2D / compute / D2H over three streams)
must return exactly what the plain upload -> extract -> match -> download sequence returns, in both modes, for
chunk sizes that do and do not divide the frame count."""
import os

import numpy as np
import pytest

from conftest import DATA

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("mode", ["reference", "orb"])
def test_process_equals_stepwise(gpu_ctx, mode):
    import torch

    import slam_cin0051_b200 as s
    from slam_cin0051_b200.synth import make_sequence
    sfx = "_orb" if mode == "orb" else ""
    det = s.FeatureDetector(os.path.join(DATA, f"feature_detector{sfx}.yml"), gpu_ctx)
    mat = s.FeatureMatcher(os.path.join(DATA, f"feature_matcher{sfx}.yml"), gpu_ctx)
    with_kp = mode == "reference"
    n, rows, cols, cap = 11, 240, 333, 2048
    frames = make_sequence(rows, cols, n, pitch_px=14, seed=21)
    ref = s.FrameSequence(rows, cols, n, desc_bytes=32, max_keypoints=cap, context=gpu_ctx)
    ref.upload(frames)
    ref.extract(det)
    ref.match_consecutive(mat, with_keypoints=with_kp)
    want_counts = ref.counts()
    assert (want_counts[:, 3] == 0).all() and want_counts[:, 0].min() > 50
    want = [(ref.frame(f), ref.matches(f) if f < n - 1 else None) for f in range(n)]
    h_frames = torch.empty((n, rows, cols), dtype=torch.uint8, pin_memory=True)
    h_frames.numpy()[:] = frames
    for chunk in (4, 11, 64, 1):
        seq = s.FrameSequence(rows, cols, n, desc_bytes=32, max_keypoints=cap, context=gpu_ctx)
        h_kps = torch.zeros((n, cap, 5), dtype=torch.float32, pin_memory=True)
        h_desc = torch.zeros((n, cap, 32), dtype=torch.uint8, pin_memory=True)
        h_m = torch.zeros((n, cap, 3), dtype=torch.int32, pin_memory=True)
        h_c = torch.zeros((n, 4), dtype=torch.int32, pin_memory=True)
        for rep in range(2):  # twice: the second call reuses the staging block and the events
            seq.process_ptrs(det, mat, h_frames.data_ptr(), n, chunk=chunk, with_keypoints=with_kp, kps_ptr=h_kps.data_ptr(),
                             desc_ptr=h_desc.data_ptr(), matches_ptr=h_m.data_ptr(), counts_ptr=h_c.data_ptr())
            gpu_ctx.synchronize()
            assert np.array_equal(h_c.numpy(), want_counts), (chunk, rep)
            for f in range(n):
                (wk, wd), wm = want[f]
                k = len(wk)
                assert h_kps.numpy()[f, :k].tobytes() == wk.tobytes(), (chunk, f)
                assert np.array_equal(h_desc.numpy()[f, :k], wd), (chunk, f)
                if wm is not None:
                    assert h_m.numpy()[f, :len(wm)].tobytes() == wm.tobytes(), (chunk, f)


@pytest.mark.parametrize("mode", ["reference", "orb"])
def test_sequence_with_empty_frames(gpu_ctx, mode):
    """Frames without a single corner inside a sequence: zero keypoints, zero matches for the pairs they take part in,
    no RANSAC result, and the neighbours are unaffected (the reference would throw "Empty descriptors provided." for such
    a pair; the batched path reports n_matches = 0)."""
    import slam_cin0051_b200 as s
    from slam_cin0051_b200.synth import make_sequence
    sfx = "_orb" if mode == "orb" else ""
    det = s.FeatureDetector(os.path.join(DATA, f"feature_detector{sfx}.yml"), gpu_ctx)
    mat = s.FeatureMatcher(os.path.join(DATA, f"feature_matcher{sfx}.yml"), gpu_ctx)
    frames = make_sequence(240, 333, 5, pitch_px=14, seed=4)
    frames[2] = 90  # flat
    seq = s.FrameSequence(240, 333, 5, desc_bytes=32, max_keypoints=2048, context=gpu_ctx)
    seq.upload(frames)
    seq.extract(det)
    seq.match_consecutive(mat, with_keypoints=(mode == "reference"))
    seq.essential((300.0, 300.0, 166.0, 120.0))
    c = seq.counts()
    assert c[2, 0] == 0 and c[1, 1] == 0 and c[2, 1] == 0 and (c[:, 3] == 0).all()
    assert c[0, 0] > 50 and c[0, 1] > 0 and c[3, 1] > 0
    E, mask, good, iters = seq.essential_result(1)
    assert E is None and good == 0 and len(mask) == 0
    k0, d0 = seq.frame(0)
    k0s, d0s = det.detect_and_compute(frames[0])
    assert k0.tobytes() == k0s.tobytes() and np.array_equal(d0, d0s)
    k2, d2 = seq.frame(2)
    assert len(k2) == 0
