"""ctypes binding of libslamcu.so (C ABI: include/slam/cuda/slamcu.h).

There is no CPU fallback: if the CUDA library is missing or cannot be loaded this module raises, and
every compute entry point raises when no CUDA device is present.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
# SLAMCU_LIB selects another build of the same library (the debug-bounds build libslamcu_dbg.so in tests/test_gpu_debug_bounds.py)
LIB_PATH = os.environ.get("SLAMCU_LIB") or os.path.join(HERE, "libslamcu.so")

OK, INVALID_ARGUMENT, EMPTY_INPUT, SIZE_MISMATCH, CAPACITY, CUDA_ERROR, UNSUPPORTED = range(7)

KEYPOINT_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"), ("response", "<f4")])
MATCH_DTYPE = np.dtype([("queryIdx", "<i4"), ("trainIdx", "<i4"), ("distance", "<f4")])
KNN2_DTYPE = np.dtype([("trainIdx0", "<i4"), ("trainIdx1", "<i4"), ("distance0", "<f4"), ("distance1", "<f4")])


class DetectorConfig(C.Structure):
    _fields_ = [("intensity_threshold", C.c_int32), ("contiguous_pixels_threshold", C.c_int32),
                ("non_max_suppression", C.c_int32), ("suppression_window_size", C.c_int32),
                ("patch_size", C.c_int32), ("num_brief_pairs", C.c_int32), ("n_pattern", C.c_int32),
                ("pattern", C.POINTER(C.c_int32)), ("blur_weights", C.POINTER(C.c_double)),
                ("mode", C.c_int32), ("n_levels", C.c_int32), ("scale_factor", C.c_float),
                ("max_features", C.c_int32), ("fast_threshold", C.c_int32), ("orb_pattern", C.POINTER(C.c_int32))]


class MatcherConfig(C.Structure):
    _fields_ = [("distance_type", C.c_int32), ("filter_matches", C.c_int32), ("good_matches_count", C.c_int32),
                ("use_ratio_test", C.c_int32), ("ratio_test_threshold", C.c_float)]


# every symbol include/slam/cuda/slamcu.h declares: name -> (restype, argtypes)
_vp, _i, _ip, _u8p, _f64p, _f32p = C.c_void_p, C.c_int, C.POINTER(C.c_int), C.c_void_p, C.c_void_p, C.c_void_p
SYMBOLS = {
    "slamcu_abi_version": (_i, []),
    "slamcu_status_string": (C.c_char_p, [_i]),
    "slamcu_device_count": (_i, [_ip]),
    "slamcu_create": (_i, [_i, C.POINTER(_vp)]),
    "slamcu_destroy": (None, [_vp]),
    "slamcu_last_error": (C.c_char_p, [_vp]),
    "slamcu_set_stream": (_i, [_vp, _vp]),
    "slamcu_get_stream": (_vp, [_vp]),
    "slamcu_synchronize": (_i, [_vp]),
    "slamcu_launch_count": (C.c_int64, [_vp]),
    "slamcu_alloc_pinned": (_i, [C.c_size_t, C.POINTER(_vp)]),
    "slamcu_free_pinned": (None, [_vp]),
    "slamcu_popc_peak": (_i, [_vp, C.POINTER(C.c_double)]),
    "slamcu_debug_trip_bound": (_i, [_vp]),
    "slamcu_profile_enable": (_i, [_vp, _i]),
    "slamcu_profile_read": (_i, [_vp, _i, C.c_char_p, _i, C.POINTER(C.c_double), C.POINTER(C.c_int64)]),
    "slamcu_default_brief_pattern": (_i, [_i, _i, _vp, _i, _ip]),
    "slamcu_default_blur_weights": (_i, [_vp]),
    "slamcu_detector_create": (_i, [_vp, C.POINTER(DetectorConfig), C.POINTER(_vp)]),
    "slamcu_detector_destroy": (None, [_vp]),
    "slamcu_detect": (_i, [_vp, _u8p, _i, _i, _i, _vp, _i, _ip]),
    "slamcu_compute": (_i, [_vp, _u8p, _i, _i, _i, _vp, _i, _u8p, _i]),
    "slamcu_detect_and_compute": (_i, [_vp, _u8p, _i, _i, _i, _vp, _u8p, _i, _i, _ip]),
    "slamcu_fast_corners": (_i, [_vp, _u8p, _i, _i, _i, _vp, _i, _ip]),
    "slamcu_detector_last_octaves": (_i, [_vp, _vp, _i]),
    "slamcu_orb_stage": (_i, [_vp, _i, _i, _vp, _vp, _i, _ip]),
    "slamcu_orb_level_image": (_i, [_vp, _i, _i, _u8p, _i, _ip, _ip]),
    "slamcu_sequence_octaves": (_i, [_vp, _i, _vp, _i]),
    "slamcu_gaussian_blur": (_i, [_vp, _u8p, _i, _i, _i, _u8p, _i]),
    "slamcu_matcher_create": (_i, [_vp, C.POINTER(MatcherConfig), C.POINTER(_vp)]),
    "slamcu_matcher_destroy": (None, [_vp]),
    "slamcu_match": (_i, [_vp, _u8p, _i, _i, _u8p, _i, _i, _vp, _i, _vp, _i, _vp, _i, _ip]),
    "slamcu_knn2_hamming": (_i, [_vp, _u8p, _i, _u8p, _i, _i, _vp]),
    "slamcu_matcher_set_train_slices": (_i, [_vp, _i]),
    "slamcu_sequence_create": (_i, [_vp, _i, _i, _i, _i, _i, _i, C.POINTER(_vp)]),
    "slamcu_sequence_destroy": (None, [_vp]),
    "slamcu_sequence_upload": (_i, [_vp, _i, _i, _u8p, _i]),
    "slamcu_sequence_prepare": (_i, [_vp, _i, _i, _u8p, _i, _i, _f64p, _f64p]),
    "slamcu_sequence_image": (_i, [_vp, _i, _u8p, _i]),
    "slamcu_sequence_frames_device": (_i, [_vp, C.POINTER(_vp), _ip, C.POINTER(C.c_int64)]),
    "slamcu_sequence_extract": (_i, [_vp, _vp, _i, _i]),
    "slamcu_sequence_match": (_i, [_vp, _vp, _i, _i, _i]),
    "slamcu_sequence_extract_match": (_i, [_vp, _vp, _vp, _i, _i, _i, _i]),
    "slamcu_sequence_counts": (_i, [_vp, _i, _i, _vp]),
    "slamcu_sequence_frame": (_i, [_vp, _i, _vp, _u8p, _i, _i, _ip]),
    "slamcu_sequence_matches": (_i, [_vp, _i, _vp, _i, _ip]),
    "slamcu_sequence_process": (_i, [_vp, _vp, _vp, _u8p, _i, _i, _i, _i, _vp, _u8p, _vp, _vp]),
    "slamcu_sequence_process_dense": (_i, [_vp, _vp, _vp, _u8p, _i, _i, _i, _i, _vp, _u8p, _vp, _vp, C.c_int64, C.c_int64]),
    "slamcu_sequence_counts_device": (_i, [_vp, _i, _i, _vp]),
    "slamcu_sequence_wait": (_i, [_vp]),
    "slamcu_sequence_download": (_i, [_vp, _i, _i, _vp, _u8p, _vp, _vp]),
    "slamcu_bgr_to_gray": (_i, [_vp, _u8p, _i, _i, _i, _u8p, _i]),
    "slamcu_undistort": (_i, [_vp, _u8p, _i, _i, _i, _f64p, _f64p, _u8p, _f64p]),
    "slamcu_ransac_score": (_i, [_vp, _f64p, _i, _f64p, _f64p, _i, C.c_double, _vp, _u8p]),
    "slamcu_find_essential": (_i, [_vp, _f32p, _f32p, _i, _f64p, C.c_double, C.c_double, _i, _f64p, _u8p, _ip]),
    "slamcu_estimate_pose": (_i, [_vp, _f32p, _f32p, _i, _f64p, _f64p, _u8p, _ip, _f64p, _f64p, _vp]),
    "slamcu_triangulate": (_i, [_vp, _f64p, _f64p, _f32p, _f32p, _i, _f64p, _f64p]),
    "slamcu_pnp_sample_indices": (_i, [C.c_uint32, _i, _i, _vp]),
    "slamcu_pnp_ransac": (_i, [_vp, _f64p, _f64p, _i, _vp, _i, _f64p, C.c_double, _vp, _f64p]),
    "slamcu_fivept_solve": (_i, [_vp, _f64p, _f64p, _i, _f64p, _vp]),
    "slamcu_sequence_essential": (_i, [_vp, _i, _i, _f64p, C.c_double, C.c_double, _i]),
    "slamcu_sequence_essential_read": (_i, [_vp, _i, _f64p, _ip, _ip, _u8p, _i, _ip]),
}

_lib = None


def load() -> C.CDLL:
    """Loads libslamcu.so and binds every declared symbol.  Raises if the library is absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} is missing: run `python -m slam_cin0051_b200.build` (there is no CPU fallback)")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)  # AttributeError if the .so does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def ptr(a):
    return None if a is None else a.ctypes.data


class SlamcuError(RuntimeError):
    def __init__(self, status, message):
        super().__init__(message)
        self.status = status


def raise_for(status: int, ctx_handle=None, what: str = ""):
    """Translates a status into the exception type the reference throws for the same condition."""
    if status == OK:
        return
    lib = load()
    msg = ""
    if ctx_handle:
        raw = lib.slamcu_last_error(ctx_handle)
        msg = raw.decode() if raw else ""
    if not msg:
        msg = lib.slamcu_status_string(status).decode()
    if what:
        msg = f"{what}: {msg}"
    if status == EMPTY_INPUT:
        raise ValueError(msg)  # std::invalid_argument in the reference
    raise SlamcuError(status, msg)  # std::runtime_error in the reference


class Context:
    """One per process / GPU.  Wraps slamcu_context."""
    _default = {}

    def __init__(self, device: int = 0):
        lib = load()
        n = C.c_int(0)
        if lib.slamcu_device_count(C.byref(n)) != OK or n.value <= 0:
            raise SlamcuError(CUDA_ERROR, "no CUDA device available: slam_cin0051_b200 has no CPU fallback")
        h = C.c_void_p()
        raise_for(lib.slamcu_create(device, C.byref(h)), None, "slamcu_create")
        self.handle = h
        self.device = device
        self.lib = lib

    @classmethod
    def default(cls, device: int = 0) -> "Context":
        if device not in cls._default:
            cls._default[device] = cls(device)
        return cls._default[device]

    def set_stream(self, cuda_stream_ptr: int | None):
        raise_for(self.lib.slamcu_set_stream(self.handle, C.c_void_p(cuda_stream_ptr or 0)), self.handle)

    def synchronize(self):
        raise_for(self.lib.slamcu_synchronize(self.handle), self.handle)

    @property
    def launch_count(self) -> int:
        return int(self.lib.slamcu_launch_count(self.handle))

    def popc_peak(self) -> float:
        v = C.c_double(0)
        raise_for(self.lib.slamcu_popc_peak(self.handle, C.byref(v)), self.handle)
        return v.value

    def profile_enable(self, on: bool = True):
        raise_for(self.lib.slamcu_profile_enable(self.handle, 1 if on else 0), self.handle)

    def profile_read(self) -> dict:
        """{kernel name: (total_ms, launches)} accumulated since profile_enable()."""
        out, i = {}, 0
        while True:
            name = C.create_string_buffer(64)
            ms, cnt = C.c_double(0), C.c_int64(0)
            if self.lib.slamcu_profile_read(self.handle, i, name, 64, C.byref(ms), C.byref(cnt)) != OK:
                break
            out[name.value.decode()] = (ms.value, cnt.value)
            i += 1
        return out

    def check(self, status, what=""):
        raise_for(status, self.handle, what)
