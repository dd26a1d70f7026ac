"""Host-side mirror of include/slam/common/common.hpp: Hamming distance and slam::Camera."""
from __future__ import annotations

import os

import numpy as np

from ._lib import Context
from .config import read_yaml


class Camera:
    """slam::Camera (common.hpp:67-190): intrinsics K<i>, distortion D<i>, ImageSize from an OpenCV YAML."""

    def __init__(self, config_path, camera_index: int = 0, context: Context | None = None):
        try:
            cfg = read_yaml(config_path)
        except RuntimeError:
            raise RuntimeError("Could not open calibration file: " + os.fspath(config_path))
        k_key, d_key = f"K{camera_index}", f"D{camera_index}"
        if k_key not in cfg or d_key not in cfg:
            raise RuntimeError(f"Could not find keys {k_key} or {d_key} in file.")
        self.K = np.asarray(cfg[k_key], np.float64).reshape(3, 3)
        self.D = np.asarray(cfg[d_key], np.float64).reshape(-1)
        size = cfg.get("ImageSize", [0, 0])
        self.image_size = (int(size[0]), int(size[1]))  # (width, height)
        self.fx, self.fy, self.cx, self.cy = self.K[0, 0], self.K[1, 1], self.K[0, 2], self.K[1, 2]
        d = list(self.D) + [0.0] * 5
        self.k1, self.k2, self.p1, self.p2, self.k3 = d[:5]  # k3 is loaded but unused (common.hpp:113,151-154)
        self._ctx = context

    @property
    def ctx(self) -> Context:
        if self._ctx is None:
            self._ctx = Context.default()
        return self._ctx

    def get_intrinsic_matrix(self) -> np.ndarray:
        return self.K

    def get_distortion_coefficients(self) -> np.ndarray:
        return self.D

    def _check(self, raw):
        img = np.asarray(raw)
        if img.size == 0:
            raise RuntimeError("Input image is empty.")
        if img.ndim != 2 or img.dtype != np.uint8:
            raise RuntimeError("Input image must be 8-bit grayscale.")
        if img.shape[0] != self.image_size[1] or img.shape[1] != self.image_size[0]:
            raise RuntimeError("Input image size does not match camera image size.")
        return np.ascontiguousarray(img)

    def _run(self, img, want_u8, want_f64):
        rows, cols = img.shape
        K4 = np.array([self.fx, self.fy, self.cx, self.cy], np.float64)
        D4 = np.array([self.k1, self.k2, self.p1, self.p2], np.float64)
        u8 = np.zeros((rows, cols), np.uint8) if want_u8 else None
        f64 = np.zeros((rows, cols), np.float64) if want_f64 else None
        self.ctx.check(self.ctx.lib.slamcu_undistort(self.ctx.handle, img.ctypes.data, rows, cols, img.strides[0],
                                                     K4.ctypes.data, D4.ctypes.data, u8.ctypes.data if want_u8 else None,
                                                     f64.ctypes.data if want_f64 else None))
        return u8, f64

    def undistort_image(self, raw_image) -> np.ndarray:
        """Camera::undistortImage (common.hpp:127-173): float64 image in [0, 1]."""
        return self._run(self._check(raw_image), False, True)[1]

    def undistort_image_u8(self, raw_image) -> np.ndarray:
        """The same gather, emitted as the uint8 image the detector consumes."""
        return self._run(self._check(raw_image), True, False)[0]

    undistortImage = undistort_image


def bgr_to_gray(bgr, context: Context | None = None) -> np.ndarray:
    """cv::cvtColor(frame, COLOR_BGR2GRAY) as called by Preprocessor::yield (preprocessor.cpp:136)."""
    ctx = context or Context.default()
    a = np.ascontiguousarray(bgr)
    if a.ndim != 3 or a.shape[2] != 3 or a.dtype != np.uint8:
        raise RuntimeError("expected an HxWx3 uint8 BGR image")
    out = np.zeros(a.shape[:2], np.uint8)
    ctx.check(ctx.lib.slamcu_bgr_to_gray(ctx.handle, a.ctypes.data, a.shape[0], a.shape[1], a.strides[0], out.ctypes.data,
                                         out.strides[0]))
    return out
