"""Host-side mirror of the reference's src/frontend interfaces, on top of the C ABI.

Same class and method names, argument meaning and error behaviour as
  slam::FeatureDetector  (include/slam/frontend/feature_detector.hpp:47-192)
  slam::FeatureMatcher   (include/slam/frontend/feature_matcher.hpp:38-87)
Images are row-major uint8 numpy arrays (EigenGrayMatrix), keypoints are structured arrays with the
fields of slam::Keypoint, descriptors are (N, NumBRIEFPairs/8) uint8 (DescriptorMatrix), matches are
structured arrays with the fields of slam::Match.  std::runtime_error -> RuntimeError,
std::invalid_argument -> ValueError.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import _lib
from ._lib import KEYPOINT_DTYPE, KNN2_DTYPE, MATCH_DTYPE, Context
from .config import get_float, get_int, get_str, read_yaml


def _gray(image) -> np.ndarray:
    a = np.asarray(image)
    if a.dtype != np.uint8 or a.ndim != 2:
        raise RuntimeError("image must be a 2-D uint8 array (EigenGrayMatrix)")
    if not a.flags.c_contiguous:
        a = np.ascontiguousarray(a)
    return a


class FeatureDetector:
    """slam::FeatureDetector.  `config` is a path to the YAML file the reference reads, or a dict with the same keys."""

    def __init__(self, config, context: Context | None = None):
        if isinstance(config, dict):
            cfg = dict(config)
        else:
            try:
                cfg = read_yaml(config)
            except RuntimeError:
                raise RuntimeError("Could not open feature detector file: " + os.fspath(config))
        self.ctx = context or Context.default()
        lib = self.ctx.lib
        self.intensity_threshold = get_int(cfg, "IntensityThreshold")
        self.contiguous_pixels_threshold = get_int(cfg, "ContiguousPixelsThreshold")
        self.non_max_suppression = get_int(cfg, "NonMaxSuppression")
        self.suppression_window_size = get_int(cfg, "SuppressionWindowSize")
        self.patch_size = get_int(cfg, "PatchSize")
        self.num_brief_pairs = get_int(cfg, "NumBRIEFPairs")
        # range checks in the reference's order and wording (feature_detector.hpp:60-93)
        if not 0 <= self.intensity_threshold <= 255:
            raise RuntimeError("Intensity threshold must be in the range [0, 255].")
        if not 0 <= self.contiguous_pixels_threshold <= 16:
            raise RuntimeError("Contiguous pixels threshold must be in the range [0, 16].")
        if self.non_max_suppression not in (0, 1):
            raise RuntimeError("Non-max suppression must be either 0 (false) or 1 (true).")
        if self.suppression_window_size <= 0:
            raise RuntimeError("Suppression window size must be a positive integer.")
        if self.patch_size <= 0 or self.patch_size % 2 == 0:
            raise RuntimeError("Patch size must be a positive odd integer.")
        if self.num_brief_pairs <= 0 or self.num_brief_pairs % 8 != 0:
            raise RuntimeError("Number of BRIEF pairs must be a positive multiple of 8.")
        # constructor-time host tables (generateBRIEFPattern, feature_detector.hpp:98)
        pat = np.zeros((self.num_brief_pairs, 4), np.int32)
        n = C.c_int(0)
        self.ctx.check(lib.slamcu_default_brief_pattern(self.patch_size, self.num_brief_pairs, pat.ctypes.data,
                                                        self.num_brief_pairs, C.byref(n)))
        self.brief_pattern = pat[: n.value].copy()
        self.blur_weights = np.zeros(25, np.float64)
        self.ctx.check(lib.slamcu_default_blur_weights(self.blur_weights.ctypes.data))
        c = _lib.DetectorConfig()
        c.intensity_threshold = self.intensity_threshold
        c.contiguous_pixels_threshold = self.contiguous_pixels_threshold
        c.non_max_suppression = self.non_max_suppression
        c.suppression_window_size = self.suppression_window_size
        c.patch_size = self.patch_size
        c.num_brief_pairs = self.num_brief_pairs
        c.n_pattern = len(self.brief_pattern)
        c.pattern = self.brief_pattern.ctypes.data_as(C.POINTER(C.c_int32))
        c.blur_weights = self.blur_weights.ctypes.data_as(C.POINTER(C.c_double))
        # Opt-in OpenCV-ORB-compatible mode: present only when the YAML carries keys the reference does not have.
        self.orb_mode = any(k in cfg for k in ("NumLevels", "MaxFeatures", "ScaleFactor"))
        c.mode = 0
        if self.orb_mode:
            self.num_levels = get_int(cfg, "NumLevels") or 8
            self.scale_factor = float(np.float32(get_float(cfg, "ScaleFactor") or 1.2))
            self.max_features = get_int(cfg, "MaxFeatures") or 2000
            if self.patch_size != 31 or self.num_brief_pairs != 256:
                raise RuntimeError("ORB mode requires PatchSize 31 and NumBRIEFPairs 256.")
            c.mode = 1
            c.n_levels = self.num_levels
            c.scale_factor = self.scale_factor
            c.max_features = self.max_features
            c.fast_threshold = get_int(cfg, "FastThreshold") or self.intensity_threshold
            c.orb_pattern = None  # the library's built-in copy of OpenCV's bit_pattern_31_
        h = C.c_void_p()
        self.ctx.check(lib.slamcu_detector_create(self.ctx.handle, C.byref(c), C.byref(h)))
        self.handle = h

    def __del__(self):
        try:
            if getattr(self, "handle", None):
                self.ctx.lib.slamcu_detector_destroy(self.handle)
                self.handle = None
        except Exception:
            pass

    @property
    def descriptor_bytes(self) -> int:
        return self.num_brief_pairs // 8

    def _run(self, fn, image, with_desc):
        img = _gray(image)
        rows, cols = img.shape
        cap = max(4096, rows * cols // 16)
        lib = self.ctx.lib
        while True:
            kps = np.zeros(cap, KEYPOINT_DTYPE)
            n = C.c_int(0)
            if with_desc:
                desc = np.zeros((cap, self.descriptor_bytes), np.uint8)
                st = fn(self.handle, img.ctypes.data, rows, cols, img.strides[0], kps.ctypes.data, desc.ctypes.data,
                        self.descriptor_bytes, cap, C.byref(n))
            else:
                desc = None
                st = fn(self.handle, img.ctypes.data, rows, cols, img.strides[0], kps.ctypes.data, cap, C.byref(n))
            if st == _lib.CAPACITY and n.value > cap:
                cap = n.value
                continue
            self.ctx.check(st)
            return (kps[: n.value].copy(), desc[: n.value].copy()) if with_desc else kps[: n.value].copy()

    def last_octaves(self, n: int) -> np.ndarray:
        """ORB mode: cv::KeyPoint::octave of the n keypoints returned by the last single-frame call."""
        out = np.zeros(max(n, 1), np.int32)
        self.ctx.check(self.ctx.lib.slamcu_detector_last_octaves(self.handle, out.ctypes.data, len(out)))
        return out[:n].copy()

    def orb_stage(self, stage: int, level: int):
        """ORB mode stage probe of the last single-frame call: (x, y, value) arrays in level coordinates."""
        n = C.c_int(0)
        lib = self.ctx.lib
        st = lib.slamcu_orb_stage(self.handle, stage, level, None, None, 0, C.byref(n))
        if st not in (_lib.OK, _lib.CAPACITY):
            self.ctx.check(st)
        xy = np.zeros(max(n.value, 1), np.uint32)
        val = np.zeros(max(n.value, 1), np.float32)
        self.ctx.check(lib.slamcu_orb_stage(self.handle, stage, level, xy.ctypes.data, val.ctypes.data, len(xy), C.byref(n)))
        xy, val = xy[: n.value], val[: n.value]
        return (xy & 0xFFFF).astype(np.int32), (xy >> 16).astype(np.int32), val.copy()

    def orb_level_image(self, level: int, blurred: bool = False) -> np.ndarray:
        """ORB mode: pyramid level (optionally its 7x7 blurred copy) of the last single-frame call."""
        r, c = C.c_int(0), C.c_int(0)
        lib = self.ctx.lib
        self.ctx.check(lib.slamcu_orb_level_image(self.handle, level, int(blurred), None, 0, C.byref(r), C.byref(c)))
        out = np.zeros((r.value, c.value), np.uint8)
        self.ctx.check(lib.slamcu_orb_level_image(self.handle, level, int(blurred), out.ctypes.data, out.strides[0],
                                                  C.byref(r), C.byref(c)))
        return out

    def detect(self, image) -> np.ndarray:
        """FeatureDetector::detect (feature_detector.cpp:8-18)."""
        return self._run(self.ctx.lib.slamcu_detect, image, False)

    def fast_corners(self, image) -> np.ndarray:
        """Raster-order FAST corners with SAD score in `response` (stage probe)."""
        return self._run(self.ctx.lib.slamcu_fast_corners, image, False)

    def compute(self, image, keypoints):
        """FeatureDetector::compute (feature_detector.cpp:20-47): returns (keypoints with angle, descriptors)."""
        img = _gray(image)
        kps = np.array(keypoints, dtype=KEYPOINT_DTYPE, copy=True)
        if len(kps) == 0:
            return kps, np.zeros((0, 0), np.uint8)  # DescriptorMatrix(0, 0)
        desc = np.zeros((len(kps), self.descriptor_bytes), np.uint8)
        self.ctx.check(self.ctx.lib.slamcu_compute(self.handle, img.ctypes.data, img.shape[0], img.shape[1],
                                                   img.strides[0], kps.ctypes.data, len(kps), desc.ctypes.data,
                                                   self.descriptor_bytes))
        return kps, desc

    def detect_and_compute(self, image):
        """FeatureDetector::detectAndCompute (feature_detector.cpp:49-54)."""
        kps, desc = self._run(self.ctx.lib.slamcu_detect_and_compute, image, True)
        if len(kps) == 0:
            desc = np.zeros((0, 0), np.uint8)
        return kps, desc

    detectAndCompute = detect_and_compute

    def gaussian_blur(self, image) -> np.ndarray:
        """FeatureDetector::gaussianBlur(image, 5, 1.0) (feature_detector.cpp:315-364)."""
        img = _gray(image)
        out = np.zeros_like(img)
        self.ctx.check(self.ctx.lib.slamcu_gaussian_blur(self.handle, img.ctypes.data, img.shape[0], img.shape[1],
                                                         img.strides[0], out.ctypes.data, out.strides[0]))
        return out


class FeatureMatcher:
    """slam::FeatureMatcher (feature_matcher.cpp:18-111)."""

    def __init__(self, config, context: Context | None = None):
        if isinstance(config, dict):
            cfg = dict(config)
        else:
            try:
                cfg = read_yaml(config)
            except RuntimeError:
                raise RuntimeError("Could not open feature matcher config file: " + os.fspath(config))
        self.ctx = context or Context.default()
        dt = get_str(cfg, "DistanceType")
        if dt == "HAMMING":
            self.distance_type = 0
        elif dt == "L2":
            self.distance_type = 1
        else:
            raise RuntimeError("Invalid distance type. Must be 'HAMMING' or 'L2'.")
        self.filter_matches = get_int(cfg, "FilterMatches")
        if self.filter_matches not in (0, 1):
            raise RuntimeError("FilterMatches must be either 0 (false) or 1 (true).")
        self.good_matches_count = get_int(cfg, "GoodMatchesCount")
        if self.filter_matches and self.good_matches_count <= 0:
            raise RuntimeError("GoodMatchesCount must be positive when filtering is enabled.")
        self.use_ratio_test = get_int(cfg, "UseRatioTest")
        if self.use_ratio_test not in (0, 1):
            raise RuntimeError("UseRatioTest must be either 0 (false) or 1 (true).")
        self.ratio_test_threshold = float(np.float32(get_float(cfg, "RatioTestThreshold")))
        if self.ratio_test_threshold < 0.0 or self.ratio_test_threshold > 1.0:
            raise RuntimeError("RatioTestThreshold must be in the range [0, 1].")
        c = _lib.MatcherConfig(self.distance_type, self.filter_matches, self.good_matches_count, self.use_ratio_test,
                               self.ratio_test_threshold)
        h = C.c_void_p()
        self.ctx.check(self.ctx.lib.slamcu_matcher_create(self.ctx.handle, C.byref(c), C.byref(h)))
        self.handle = h

    def __del__(self):
        try:
            if getattr(self, "handle", None):
                self.ctx.lib.slamcu_matcher_destroy(self.handle)
                self.handle = None
        except Exception:
            pass

    @staticmethod
    def _desc(d):
        a = np.asarray(d)
        if a.ndim != 2:
            a = a.reshape(len(a), -1) if a.size else np.zeros((0, 0), np.uint8)
        if a.dtype != np.uint8:
            raise RuntimeError("DescriptorMatrix must be uint8")
        return np.ascontiguousarray(a)

    def match(self, descriptors1, descriptors2, keypoints1=None, keypoints2=None) -> np.ndarray:
        """FeatureMatcher::match (feature_matcher.cpp:71-95).  Returns a MATCH_DTYPE array."""
        d1, d2 = self._desc(descriptors1), self._desc(descriptors2)
        k1 = np.ascontiguousarray(keypoints1, KEYPOINT_DTYPE) if keypoints1 is not None else np.zeros(0, KEYPOINT_DTYPE)
        k2 = np.ascontiguousarray(keypoints2, KEYPOINT_DTYPE) if keypoints2 is not None else np.zeros(0, KEYPOINT_DTYPE)
        n1, n2 = d1.shape[0], d2.shape[0]
        out = np.zeros(max(n1, 1), MATCH_DTYPE)
        n = C.c_int(0)
        st = self.ctx.lib.slamcu_match(self.handle, d1.ctypes.data if n1 else None, n1, d1.shape[1] if n1 else 0,
                                       d2.ctypes.data if n2 else None, n2, d2.shape[1] if n2 else 0,
                                       k1.ctypes.data if len(k1) else None, len(k1),
                                       k2.ctypes.data if len(k2) else None, len(k2), out.ctypes.data, len(out),
                                       C.byref(n))
        self.ctx.check(st)
        return out[: n.value].copy()

    def knn2(self, descriptors1, descriptors2) -> np.ndarray:
        """All-pairs Hamming k=2 search (best and second by (distance, trainIdx)); KNN2_DTYPE per query."""
        d1, d2 = self._desc(descriptors1), self._desc(descriptors2)
        if d1.shape[0] == 0 or d2.shape[0] == 0:
            raise ValueError("Empty descriptors provided.")
        if d1.shape[1] != d2.shape[1]:
            raise RuntimeError("Descriptor dimensions must match.")
        out = np.zeros(d1.shape[0], KNN2_DTYPE)
        self.ctx.check(self.ctx.lib.slamcu_knn2_hamming(self.handle, d1.ctypes.data, d1.shape[0], d2.ctypes.data,
                                                        d2.shape[0], d1.shape[1], out.ctypes.data))
        return out

    def set_train_slices(self, n_slices: int = 0):
        """Tuning / test knob (slamcu_matcher_set_train_slices): 0 = automatic; results do not depend on it."""
        self.ctx.check(self.ctx.lib.slamcu_matcher_set_train_slices(self.handle, int(n_slices)))
