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
