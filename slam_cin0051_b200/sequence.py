"""Device-resident frame sequences: the batched / asynchronous path of the frontend.

A FrameSequence keeps up to `max_frames` frames of one size in HBM with every intermediate of the
pipeline.  extract() is FeatureDetector::detectAndCompute on a range of frames; match_consecutive() is
FeatureMatcher::match(frame f, frame f+1) for a range of pairs.  Frames are independent and pairs are
independent, so a sequence shards across GPUs by contiguous frame range with no data-path collective.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from ._lib import KEYPOINT_DTYPE, MATCH_DTYPE, Context
from .frontend import FeatureDetector, FeatureMatcher


class FrameSequence:
    def __init__(self, rows: int, cols: int, max_frames: int, desc_bytes: int = 32, max_raw_corners: int = 0,
                 max_keypoints: int = 0, context: Context | None = None):
        self.ctx = context or Context.default()
        h = C.c_void_p()
        self.ctx.check(self.ctx.lib.slamcu_sequence_create(self.ctx.handle, rows, cols, max_frames, max_raw_corners,
                                                           max_keypoints, desc_bytes, C.byref(h)))
        self.handle = h
        self.rows, self.cols, self.max_frames, self.desc_bytes = rows, cols, max_frames, desc_bytes
        dptr, pitch, fb = C.c_void_p(), C.c_int(), C.c_int64()
        self.ctx.check(self.ctx.lib.slamcu_sequence_frames_device(h, C.byref(dptr), C.byref(pitch), C.byref(fb)))
        self.device_ptr, self.pitch, self.frame_bytes = dptr.value, pitch.value, fb.value

    def __del__(self):
        try:
            if getattr(self, "handle", None):
                self.ctx.lib.slamcu_sequence_destroy(self.handle)
                self.handle = None
        except Exception:
            pass

    def upload(self, frames: np.ndarray, first: int = 0):
        """frames: (n, rows, cols) uint8 host array (pinned memory makes the copy asynchronous)."""
        a = np.asarray(frames)
        if a.ndim == 2:
            a = a[None]
        if a.dtype != np.uint8 or a.shape[1:] != (self.rows, self.cols) or not a.flags.c_contiguous:
            raise RuntimeError("frames must be a C-contiguous (n, rows, cols) uint8 array")
        self.ctx.check(self.ctx.lib.slamcu_sequence_upload(self.handle, first, a.shape[0], a.ctypes.data, self.cols))

    def prepare(self, frames: np.ndarray, camera=None, first: int = 0):
        """Preprocessor::yield for a batch: frames (n, rows, cols) gray or (n, rows, cols, 3) BGR uint8; BGR2GRAY and, with a
        Camera, undistortImage, written straight into the frame store."""
        a = np.ascontiguousarray(frames)
        if a.dtype != np.uint8 or a.ndim not in (3, 4) or a.shape[1:3] != (self.rows, self.cols) or (a.ndim == 4 and a.shape[3] != 3):
            raise RuntimeError("frames must be (n, rows, cols) or (n, rows, cols, 3) uint8")
        ch = 3 if a.ndim == 4 else 1
        K4 = D4 = None
        if camera is not None:
            K4 = np.array([camera.fx, camera.fy, camera.cx, camera.cy], np.float64)
            D4 = np.array([camera.k1, camera.k2, camera.p1, camera.p2], np.float64)
        self.ctx.check(self.ctx.lib.slamcu_sequence_prepare(self.handle, first, a.shape[0], a.ctypes.data, ch, self.cols * ch,
                                                            K4.ctypes.data if K4 is not None else None,
                                                            D4.ctypes.data if D4 is not None else None))
        self.ctx.synchronize()  # the host array may go away

    def image(self, f: int) -> np.ndarray:
        """The prepared 8-bit frame f as the detector sees it (device -> host)."""
        out = np.zeros((self.rows, self.cols), np.uint8)
        self.ctx.check(self.ctx.lib.slamcu_sequence_image(self.handle, f, out.ctypes.data, self.cols))
        return out

    def upload_ptr(self, host_ptr: int, n: int, first: int = 0):
        self.ctx.check(self.ctx.lib.slamcu_sequence_upload(self.handle, first, n, C.c_void_p(host_ptr), self.cols))

    def extract(self, detector: FeatureDetector, first: int = 0, n: int | None = None):
        n = self.max_frames - first if n is None else n
        self.ctx.check(self.ctx.lib.slamcu_sequence_extract(self.handle, detector.handle, first, n))

    def match_consecutive(self, matcher: FeatureMatcher, first: int = 0, n_pairs: int | None = None,
                          with_keypoints: bool = True):
        n_pairs = self.max_frames - 1 - first if n_pairs is None else n_pairs
        self.ctx.check(self.ctx.lib.slamcu_sequence_match(self.handle, matcher.handle, first, n_pairs,
                                                          1 if with_keypoints else 0))

    def extract_match(self, detector: FeatureDetector, matcher: FeatureMatcher, first: int = 0, n: int | None = None,
                      with_keypoints: bool = True, chunk: int = 0):
        """extract() on frames [first, first + n) and match_consecutive() on the n - 1 pairs inside, chunked over the context's two
        compute lanes (the matcher of one chunk overlaps the extraction of the next); same results, asynchronous."""
        n = self.max_frames - first if n is None else n
        self.ctx.check(self.ctx.lib.slamcu_sequence_extract_match(self.handle, detector.handle, matcher.handle, first, n,
                                                                  1 if with_keypoints else 0, chunk))

    def counts(self, first: int = 0, n: int | None = None) -> np.ndarray:
        """(n, 4) int32: n_keypoints, n_matches (pair f,f+1), n_raw_corners, status.  Synchronises."""
        n = self.max_frames - first if n is None else n
        out = np.zeros((n, 4), np.int32)
        self.ctx.check(self.ctx.lib.slamcu_sequence_counts(self.handle, first, n, out.ctypes.data))
        return out

    def frame(self, f: int):
        cap = 4096
        while True:
            kps = np.zeros(cap, KEYPOINT_DTYPE)
            desc = np.zeros((cap, self.desc_bytes), np.uint8)
            n = C.c_int(0)
            st = self.ctx.lib.slamcu_sequence_frame(self.handle, f, kps.ctypes.data, desc.ctypes.data, self.desc_bytes,
                                                    cap, C.byref(n))
            if st == 4 and n.value > cap:
                cap = n.value
                continue
            self.ctx.check(st)
            return kps[: n.value].copy(), desc[: n.value].copy()

    def octaves(self, f: int, n: int) -> np.ndarray:
        out = np.zeros(max(n, 1), np.int32)
        self.ctx.check(self.ctx.lib.slamcu_sequence_octaves(self.handle, f, out.ctypes.data, len(out)))
        return out[:n].copy()

    def matches(self, f: int) -> np.ndarray:
        cap = 4096
        while True:
            out = np.zeros(cap, MATCH_DTYPE)
            n = C.c_int(0)
            st = self.ctx.lib.slamcu_sequence_matches(self.handle, f, out.ctypes.data, cap, C.byref(n))
            if st == 4 and n.value > cap:
                cap = n.value
                continue
            self.ctx.check(st)
            return out[: n.value].copy()

    def process_ptrs(self, detector: FeatureDetector, matcher: FeatureMatcher, host_ptr: int, n: int, chunk: int = 64,
                     with_keypoints: bool = True, kps_ptr=None, desc_ptr=None, matches_ptr=None, counts_ptr=None):
        """upload -> extract -> match(f, f+1) -> download for n host frames (dense rows), pipelined by chunks over the
        copy engines; asynchronous (Context.synchronize() waits for the downloads)."""
        self.ctx.check(self.ctx.lib.slamcu_sequence_process(
            self.handle, detector.handle, matcher.handle, C.c_void_p(host_ptr), self.cols, n, chunk,
            1 if with_keypoints else 0, C.c_void_p(kps_ptr or 0), C.c_void_p(desc_ptr or 0), C.c_void_p(matches_ptr or 0),
            C.c_void_p(counts_ptr or 0)))

    def process_dense_ptrs(self, detector: FeatureDetector, matcher: FeatureMatcher, host_ptr: int, n: int, chunk: int = 64,
                           with_keypoints: bool = True, kps_ptr=None, desc_ptr=None, matches_ptr=None, counts_ptr=None,
                           kp_capacity: int = 0, match_capacity: int = 0):
        """process_ptrs with dense outputs (slamcu_sequence_process_dense): rows of frame f start at the running sums of the
        counts; the buffers must be page-locked (torch pin_memory / slamcu_alloc_pinned)."""
        self.ctx.check(self.ctx.lib.slamcu_sequence_process_dense(
            self.handle, detector.handle, matcher.handle, C.c_void_p(host_ptr), self.cols, n, chunk,
            1 if with_keypoints else 0, C.c_void_p(kps_ptr or 0), C.c_void_p(desc_ptr or 0), C.c_void_p(matches_ptr or 0),
            C.c_void_p(counts_ptr or 0), kp_capacity, match_capacity))

    def counts_device(self, device_ptr: int, first: int = 0, n: int | None = None):
        """Packs the per-frame counts into a device int32[n][4] array (e.g. a torch tensor) on the context's stream."""
        n = self.max_frames - first if n is None else n
        self.ctx.check(self.ctx.lib.slamcu_sequence_counts_device(self.handle, first, n, C.c_void_p(device_ptr)))

    def essential(self, K4, first: int = 0, n_pairs: int | None = None, prob: float = 0.999, threshold: float = 1.0,
                  max_iters: int = 1000):
        """cv::findEssentialMat(RANSAC) on the matches of every pair (f, f+1) of the range; asynchronous."""
        n_pairs = self.max_frames - 1 - first if n_pairs is None else n_pairs
        k = np.ascontiguousarray(K4, np.float64)
        self.ctx.check(self.ctx.lib.slamcu_sequence_essential(self.handle, first, n_pairs, k.ctypes.data, prob, threshold, max_iters))

    def essential_result(self, pair: int):
        """(E 3x3 or None, mask uint8[n_matches], n_inliers, n_iters) of one pair; synchronises."""
        E = np.zeros(9, np.float64)
        n_in, n_it, n_pt = C.c_int(0), C.c_int(0), C.c_int(0)
        cap = 1 << 16
        mask = np.zeros(cap, np.uint8)
        self.ctx.check(self.ctx.lib.slamcu_sequence_essential_read(self.handle, pair, E.ctypes.data, C.byref(n_in), C.byref(n_it),
                                                                  mask.ctypes.data, cap, C.byref(n_pt)))
        return (E.reshape(3, 3) if n_in.value > 0 else None), mask[: n_pt.value].copy(), n_in.value, n_it.value

    def wait(self):
        """Blocks until the latest process_ptrs() call on this sequence (kernels and downloads) has finished."""
        self.ctx.check(self.ctx.lib.slamcu_sequence_wait(self.handle))

    def download_ptrs(self, first, n, kps_ptr=None, desc_ptr=None, matches_ptr=None, counts_ptr=None):
        self.ctx.check(self.ctx.lib.slamcu_sequence_download(self.handle, first, n, C.c_void_p(kps_ptr or 0),
                                                             C.c_void_p(desc_ptr or 0), C.c_void_p(matches_ptr or 0),
                                                             C.c_void_p(counts_ptr or 0)))
