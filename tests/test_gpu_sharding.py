"""Multi-GPU semantics on one device: a sequence sharded by contiguous frame range with a recomputed 1-frame halo
(slam_cin0051_b200.sharding.frame_range, the split bench.py and tools/cli/slam_bench use: SURVEY 8e) gives, concatenated,
exactly the keypoints / descriptors / matches of the unsharded run -- including the pair that straddles every range
boundary -- with one context per rank; and the dense (compacted) end-to-end outputs equal the padded ones."""
import os

import numpy as np
import pytest

from conftest import DATA

pytestmark = pytest.mark.gpu


def _frames(n, rows=240, cols=400):
    from slam_cin0051_b200.synth import make_sequence
    out = np.empty((n, rows, cols), np.uint8)
    for g in range(0, n, 16):
        out[g:g + 16] = make_sequence(rows, cols, min(16, n - g), pitch_px=14, seed=40 + g // 16)
    return out


@pytest.mark.parametrize("mode,world", [("orb", 2), ("orb", 3), ("reference", 2)])
def test_sharded_ranges_equal_the_unsharded_run(mode, world):
    import slam_cin0051_b200 as s
    from slam_cin0051_b200.sharding import frame_range
    sfx = "_orb" if mode == "orb" else ""
    n = 23  # not a multiple of the world sizes: ranges of unequal length
    frames = _frames(n)

    def run(ctx, part, n_local):
        det = s.FeatureDetector(os.path.join(DATA, f"feature_detector{sfx}.yml"), ctx)
        mat = s.FeatureMatcher(os.path.join(DATA, f"feature_matcher{sfx}.yml"), ctx)
        seq = s.FrameSequence(frames.shape[1], frames.shape[2], n_local, desc_bytes=32, max_keypoints=4096, context=ctx)
        seq.upload(np.ascontiguousarray(part))
        seq.extract(det)
        seq.match_consecutive(mat, 0, n_local - 1, with_keypoints=(mode != "orb"))
        c = seq.counts()
        assert (c[:, 3] == 0).all()
        return [seq.frame(f) for f in range(n_local)], [seq.matches(f) for f in range(n_local - 1)], c

    whole_f, whole_m, whole_c = run(s.Context(0), frames, n)
    seen_frames, seen_pairs = 0, 0
    for r in range(world):
        lo, hi, halo = frame_range(n, r, world)
        nl = hi - lo + halo
        part_f, part_m, part_c = run(s.Context(0), frames[lo:lo + nl], nl)  # its own context, like a rank on its own GPU
        for f in range(hi - lo):  # owned frames (the halo frame belongs to the next rank)
            (k, d), (wk, wd) = part_f[f], whole_f[lo + f]
            assert k.tobytes() == wk.tobytes() and np.array_equal(d, wd)
            assert part_c[f, 0] == whole_c[lo + f, 0]
            seen_frames += 1
        for p in range(nl - 1):  # owned pairs: (lo + p, lo + p + 1), the last one reaching into the halo frame
            assert part_m[p].tobytes() == whole_m[lo + p].tobytes(), f"pair {lo + p} (rank {r}) differs from the unsharded run"
            assert part_c[p, 1] == whole_c[lo + p, 1]
            seen_pairs += 1
        if halo:  # the recomputed halo frame is bit-identical to the next rank's first frame
            (k, d), (wk, wd) = part_f[nl - 1], whole_f[hi]
            assert k.tobytes() == wk.tobytes() and np.array_equal(d, wd)
    assert seen_frames == n and seen_pairs == n - 1


@pytest.mark.parametrize("mode", ["orb", "reference"])
def test_dense_outputs_equal_padded_outputs(gpu_ctx, mode):
    import torch

    import slam_cin0051_b200 as s
    sfx = "_orb" if mode == "orb" else ""
    det = s.FeatureDetector(os.path.join(DATA, f"feature_detector{sfx}.yml"), gpu_ctx)
    mat = s.FeatureMatcher(os.path.join(DATA, f"feature_matcher{sfx}.yml"), gpu_ctx)
    n, cap = 11, 4096
    frames = _frames(n)
    frames[4] = 90  # a frame without keypoints in the middle: zero-length rows in the dense layout
    seq = s.FrameSequence(frames.shape[1], frames.shape[2], n, desc_bytes=32, max_keypoints=cap, context=gpu_ctx)
    h_frames = torch.empty(frames.shape, dtype=torch.uint8, pin_memory=True)
    h_frames.numpy()[:] = frames
    pk = torch.zeros((n, cap, 5), dtype=torch.float32, pin_memory=True)
    pd = torch.zeros((n, cap, 32), dtype=torch.uint8, pin_memory=True)
    pm = torch.zeros((n, cap, 3), dtype=torch.int32, pin_memory=True)
    pc = torch.zeros((n, 4), dtype=torch.int32, pin_memory=True)
    seq.process_ptrs(det, mat, h_frames.data_ptr(), n, chunk=4, with_keypoints=(mode != "orb"), kps_ptr=pk.data_ptr(), desc_ptr=pd.data_ptr(),
                     matches_ptr=pm.data_ptr(), counts_ptr=pc.data_ptr())
    seq.wait()
    c = pc.numpy().copy()
    tk, tm = int(c[:, 0].sum()), int(c[:, 1].sum())
    assert tk > 0 and tm > 0 and c[4, 0] == 0
    for chunk in (1, 4, 64):
        dk = torch.full((tk + 7, 5), -1.0, dtype=torch.float32, pin_memory=True)
        dd = torch.full((tk + 7, 32), 255, dtype=torch.uint8, pin_memory=True)
        dm = torch.full((tm + 7, 3), -1, dtype=torch.int32, pin_memory=True)
        dc = torch.zeros((n, 4), dtype=torch.int32, pin_memory=True)
        seq.process_dense_ptrs(det, mat, h_frames.data_ptr(), n, chunk=chunk, with_keypoints=(mode != "orb"), kps_ptr=dk.data_ptr(),
                               desc_ptr=dd.data_ptr(), matches_ptr=dm.data_ptr(), counts_ptr=dc.data_ptr(), kp_capacity=tk + 7, match_capacity=tm + 7)
        seq.wait()
        assert np.array_equal(dc.numpy(), c)
        ko = np.r_[0, np.cumsum(c[:, 0])]
        mo = np.r_[0, np.cumsum(c[:, 1])]
        for f in range(n):
            assert np.array_equal(dk.numpy()[ko[f]:ko[f + 1]].view(np.uint32), pk.numpy()[f, :c[f, 0]].view(np.uint32))
            assert np.array_equal(dd.numpy()[ko[f]:ko[f + 1]], pd.numpy()[f, :c[f, 0]])
            assert np.array_equal(dm.numpy()[mo[f]:mo[f + 1]], pm.numpy()[f, :c[f, 1]])
        assert (dk.numpy()[tk:] == -1.0).all() and (dd.numpy()[tk:] == 255).all() and (dm.numpy()[tm:] == -1).all()  # nothing past the end
    # capacities that are too small are reported, not silently truncated
    dk = torch.zeros((max(tk // 2, 1), 5), dtype=torch.float32, pin_memory=True)
    seq.process_dense_ptrs(det, mat, h_frames.data_ptr(), n, chunk=4, with_keypoints=(mode != "orb"), kps_ptr=dk.data_ptr(), kp_capacity=len(dk),
                           match_capacity=0)
    with pytest.raises(RuntimeError, match="dense output overflow"):
        seq.wait()


def test_counts_device_pack(gpu_ctx):
    import torch

    import slam_cin0051_b200 as s
    det = s.FeatureDetector(os.path.join(DATA, "feature_detector_orb.yml"), gpu_ctx)
    mat = s.FeatureMatcher(os.path.join(DATA, "feature_matcher_orb.yml"), gpu_ctx)
    frames = _frames(6)
    seq = s.FrameSequence(frames.shape[1], frames.shape[2], 6, desc_bytes=32, max_keypoints=4096, context=gpu_ctx)
    seq.upload(frames)
    seq.extract(det)
    seq.match_consecutive(mat, with_keypoints=False)
    d = torch.zeros((4, 4), dtype=torch.int32, device="cuda")
    seq.counts_device(d.data_ptr(), 1, 4)
    gpu_ctx.synchronize()
    assert np.array_equal(d.cpu().numpy(), seq.counts(1, 4))
