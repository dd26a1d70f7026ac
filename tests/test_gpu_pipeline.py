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


@pytest.mark.parametrize("mode", ["orb", "reference"])
def test_two_lane_extract_match_equals_the_serial_calls(gpu_ctx, mode):
    """slamcu_sequence_extract_match (chunks alternating between the two compute lanes, the matcher of one chunk next to the
    extraction of the following one) gives bit for bit what extract() followed by match_consecutive() gives."""
    import os

    import slam_cin0051_b200 as s
    from conftest import DATA
    from slam_cin0051_b200.synth import make_sequence
    sfx = "_orb" if mode == "orb" else ""
    det = s.FeatureDetector(os.path.join(DATA, f"feature_detector{sfx}.yml"), gpu_ctx)
    mat = s.FeatureMatcher(os.path.join(DATA, f"feature_matcher{sfx}.yml"), gpu_ctx)
    n = 37
    frames = np.concatenate([make_sequence(240, 400, 16, 14, seed=50 + g) for g in range(3)])[:n]
    with_kp = mode != "orb"

    def snapshot(seq):
        c = seq.counts()
        assert (c[:, 3] == 0).all()
        return c.tobytes(), [tuple(x.tobytes() for x in seq.frame(f)) for f in range(n)], [seq.matches(f).tobytes() for f in range(n - 1)]

    a = s.FrameSequence(240, 400, n, desc_bytes=32, max_keypoints=4096, context=gpu_ctx)
    a.upload(frames)
    a.extract(det)
    a.match_consecutive(mat, with_keypoints=with_kp)
    want = snapshot(a)
    for chunk in (0, 1, 5, 16, 36, 64):
        b = s.FrameSequence(240, 400, n, desc_bytes=32, max_keypoints=4096, context=gpu_ctx)
        b.upload(frames)
        for _ in range(2):  # twice: the second call reuses events and lanes
            b.extract_match(det, mat, 0, n, with_keypoints=with_kp, chunk=chunk)
        assert snapshot(b) == want, chunk
    # a sub-range
    b = s.FrameSequence(240, 400, n, desc_bytes=32, max_keypoints=4096, context=gpu_ctx)
    b.upload(frames)
    b.extract_match(det, mat, 5, 20, with_keypoints=with_kp, chunk=6)
    for f in range(5, 25):
        assert tuple(x.tobytes() for x in b.frame(f)) == want[1][f]
    for f in range(5, 24):
        assert b.matches(f).tobytes() == want[2][f]


def test_two_lanes_under_load_equal_one_lane(gpu_ctx):
    """Regression: 400 headline-sized frames in small chunks, so that the last wave of one chunk's tensor-core matcher shares its
    SMs with the other lane's extractor (a shared-memory slot reuse race showed up exactly there as a handful of wrong matches):
    per-pair match counts and the matches themselves equal the one-lane run, three times in a row."""
    import os

    import slam_cin0051_b200 as s
    from conftest import DATA
    from slam_cin0051_b200.synth import make_sequence
    det = s.FeatureDetector(os.path.join(DATA, "feature_detector_orb.yml"), gpu_ctx)
    mat = s.FeatureMatcher(os.path.join(DATA, "feature_matcher_orb.yml"), gpu_ctx)
    n = 400
    frames = np.concatenate([make_sequence(376, 1241, 50, 14, seed=900 + g) for g in range(8)])[:n]
    a = s.FrameSequence(376, 1241, n, desc_bytes=32, max_keypoints=2560, context=gpu_ctx)
    a.upload(frames)
    a.extract_match(det, mat, 0, n, with_keypoints=False, chunk=0)
    want_counts = a.counts().copy()
    assert (want_counts[:, 3] == 0).all()
    want = [a.matches(f).tobytes() for f in range(0, n - 1, 7)]
    for chunk in (25, 50, 25):
        a.extract_match(det, mat, 0, n, with_keypoints=False, chunk=chunk)
        got = a.counts()
        assert np.array_equal(got, want_counts), (chunk, np.nonzero((got != want_counts).any(1))[0][:10])
        assert [a.matches(f).tobytes() for f in range(0, n - 1, 7)] == want, chunk
