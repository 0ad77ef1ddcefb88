"""ctypes binding of librsvio_b200.so -- every call goes through the C ABI of include/rsvio_b200.h.

Class and method names follow the reference (src/feature_tracker/feature_tracker.rs):
  StereoPatchTracker::new / process_frame / get_track_points / remove_id   (:103, :116, :188, :201)
  PatchTracker::process_frame / get_track_points / remove_id               (:43, :78, :84)
  Feature::new                                                             (:25)
"""
import ctypes as C
import os
import subprocess
from dataclasses import dataclass, field

import numpy as np

_PKG = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_PKG)                      # rs-vio_b200/
# RSVIO_LIB: another build of the same library (A/B runs of kernel variants); still the CUDA library, never a fallback
_LIB_PATH = os.environ.get("RSVIO_LIB") or os.path.join(_ROOT, "lib", "librsvio_b200.so")

OK, ERR_INVALID_ARG, ERR_CUDA, ERR_CAPACITY, ERR_STATE, ERR_ALLOC = 0, -1, -2, -3, -4, -5
TRACK_OK, FAIL_TEMPLATE, FAIL_RESIDUAL, FAIL_NONFINITE, FAIL_DIVERGED, FAIL_OOB, FAIL_FB = range(7)
FAIL_BACKWARD = 16
MODE_STEREO, MODE_MONO = 0, 1
N_STAGES = 8
STAGE_NAMES = ("upload", "pyramid", "lk_temporal", "compact", "detect", "lk_stereo", "merge", "download")

# every symbol include/rsvio_b200.h declares
ABI_SYMBOLS = (
    "rsvio_abi_version", "rsvio_status_string", "rsvio_device_count", "rsvio_config_default",
    "rsvio_host_alloc", "rsvio_host_alloc_flags", "rsvio_host_free",
    "rsvio_tracker_create", "rsvio_tracker_destroy", "rsvio_tracker_last_error", "rsvio_tracker_reset",
    "rsvio_tracker_process_stereo", "rsvio_tracker_process_mono", "rsvio_tracker_process_batch",
    "rsvio_tracker_submit_batch", "rsvio_tracker_sync", "rsvio_tracker_submit_device",
    "rsvio_tracker_num_tracks", "rsvio_tracker_get_tracks", "rsvio_tracker_result_view",
    "rsvio_tracker_get_transforms",
    "rsvio_tracker_remove_ids", "rsvio_tracker_counters", "rsvio_tracker_overflow_frames",
    "rsvio_tracker_set_profiling",
    "rsvio_tracker_stage_times",
    "rsvio_pyramid_bytes", "rsvio_pyramid_device_bytes", "rsvio_build_pyramid_u8",
    "rsvio_build_pyramid_u8_device", "rsvio_track_points", "rsvio_detect_keypoints", "rsvio_se2_exp",
    "rsvio_tracker_set_camera", "rsvio_tracker_undistorted_view", "rsvio_unproject_points",
    "rsvio_tracker_process_stereo_fmt", "rsvio_convert_to_luma8", "rsvio_b_track_points",
    "rsvio_b_pyramid_dims", "rsvio_b_build_pyramid", "rsvio_microbench_fp32",
    "rsvio_b_track_points_cost", "rsvio_b_blur", "rsvio_b_shi_tomasi_score", "rsvio_b_add_points",
    "rsvio_b_config_default", "rsvio_b_tracker_create", "rsvio_b_tracker_destroy", "rsvio_b_tracker_process_frame",
    "rsvio_b_tracker_num_features", "rsvio_b_tracker_get_features", "rsvio_b_tracker_pyramid_dims",
    "rsvio_b_tracker_get_pyramid_level", "rsvio_b_tracker_last_frame_ms",
)

B_COST_SSD, B_COST_LSSD = 0, 1

PIXEL_LUMA8, PIXEL_LUMA16, PIXEL_RGB8 = 0, 1, 2

CAMERA_NONE, CAMERA_PINHOLE_RADTAN, CAMERA_EUCM = 0, 1, 2


class RsvioError(RuntimeError):
    def __init__(self, status, msg=""):
        self.status = status
        super().__init__(f"rsvio status {status}: {msg}")


class Config(C.Structure):
    _fields_ = [("width", C.c_int32), ("height", C.c_int32), ("levels", C.c_int32), ("grid_size", C.c_uint32),
                ("max_iterations", C.c_int32), ("convergence_threshold", C.c_float), ("mode", C.c_int32),
                ("n_streams", C.c_int32), ("device", C.c_int32), ("max_tracks", C.c_int32),
                ("cuda_stream", C.c_void_p)]


class BConfig(C.Structure):
    """rsvio_b_config: crate B's FeatureTrackingConfig (feature_tracker/src/feature_tracker.rs:24-38) + size / device"""
    _fields_ = [("width", C.c_int32), ("height", C.c_int32), ("nlevels", C.c_int32), ("ratio", C.c_double),
                ("preprocessing_blur", C.c_int32), ("preprocessing_blur_sigma", C.c_float),
                ("detection_threshold", C.c_float), ("detection_min_dist", C.c_uint32), ("detection_blur", C.c_float),
                ("optical_flow_max_iter", C.c_int32), ("optical_flow_lm_lambda", C.c_float),
                ("matching_cost", C.c_int32), ("device", C.c_int32), ("max_features", C.c_int32)]


def library_path():
    return _LIB_PATH


def build_library(force=False, verbose=False):
    """Compile csrc/*.cu for sm_100a with the package Makefile (nvcc cross-compiles without a GPU)."""
    args = ["make", "-C", _ROOT, "-j8"]
    if force:
        args.insert(1, "-B")
    subprocess.run(args, check=True, stdout=None if verbose else subprocess.DEVNULL)
    return _LIB_PATH


_lib = None
_u8p = np.ctypeslib.ndpointer(np.uint8, flags="C_CONTIGUOUS")
_f32p = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")
_i32p = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")
_u32p = np.ctypeslib.ndpointer(np.uint32, flags="C_CONTIGUOUS")
_u64p = np.ctypeslib.ndpointer(np.uint64, flags="C_CONTIGUOUS")


class CameraModel(C.Structure):
    """rsvio_camera_model: pinhole-radtan = fx fy cx cy k1 k2 p1 p2 k3, EUCM = fx fy cx cy alpha beta
    (the parameter vectors of the reference's create_camera_models_from_config, datasets/mod.rs:93-163)."""
    _fields_ = [("model", C.c_int32), ("reserved", C.c_int32), ("params", C.c_double * 9)]

    @classmethod
    def pinhole_radtan(cls, intrinsics, distortion):
        d = (list(distortion) + [0.0] * 5)[:5]            # k1 k2 p1 p2 k3, missing ones 0 (mod.rs:121-125)
        return cls(CAMERA_PINHOLE_RADTAN, 0, (C.c_double * 9)(*(list(intrinsics)[:4] + d)))

    @classmethod
    def eucm(cls, intrinsics, distortion):
        return cls(CAMERA_EUCM, 0, (C.c_double * 9)(*(list(intrinsics)[:4] + list(distortion)[:2] + [0.0] * 3)))


def load_library():
    """dlopen librsvio_b200.so.  Raises (never falls back) when the CUDA extension is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_LIB_PATH):
        raise RsvioError(ERR_CUDA, f"{_LIB_PATH} is missing: build it with `make -C {_ROOT}` "
                                   "(there is no CPU fallback)")
    L = C.CDLL(_LIB_PATH)
    st = C.c_int32
    H = C.c_void_p
    L.rsvio_abi_version.restype = C.c_int32
    L.rsvio_status_string.restype = C.c_char_p
    L.rsvio_status_string.argtypes = [st]
    L.rsvio_device_count.argtypes = [C.POINTER(C.c_int32)]
    L.rsvio_config_default.restype = None
    L.rsvio_config_default.argtypes = [C.POINTER(Config)]
    L.rsvio_host_alloc.argtypes = [C.POINTER(C.c_void_p), C.c_size_t]
    L.rsvio_host_alloc_flags.argtypes = [C.POINTER(C.c_void_p), C.c_size_t, C.c_uint32]
    L.rsvio_host_free.argtypes = [C.c_void_p]
    L.rsvio_tracker_create.argtypes = [C.POINTER(Config), C.POINTER(H)]
    L.rsvio_tracker_destroy.restype = None
    L.rsvio_tracker_destroy.argtypes = [H]
    L.rsvio_tracker_last_error.restype = C.c_char_p
    L.rsvio_tracker_last_error.argtypes = [H]
    L.rsvio_tracker_reset.argtypes = [H]
    L.rsvio_tracker_process_stereo.argtypes = [H, C.c_void_p, C.c_void_p, C.c_size_t]
    L.rsvio_tracker_process_mono.argtypes = [H, C.c_void_p, C.c_size_t]
    L.rsvio_tracker_process_batch.argtypes = [H, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.c_size_t]
    L.rsvio_tracker_submit_batch.argtypes = [H, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.c_size_t]
    L.rsvio_tracker_sync.argtypes = [H]
    L.rsvio_tracker_submit_device.argtypes = [H, C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t]
    L.rsvio_tracker_num_tracks.argtypes = [H, C.c_int32, C.c_int32, C.POINTER(C.c_int32)]
    L.rsvio_tracker_get_tracks.argtypes = [H, C.c_int32, C.c_int32, _u64p, _f32p, C.c_size_t, C.POINTER(C.c_size_t)]
    L.rsvio_tracker_result_view.argtypes = [H, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(C.c_int32),
                                            C.POINTER(C.c_int32)]
    L.rsvio_tracker_get_transforms.argtypes = [H, C.c_int32, C.c_int32, _u64p, _f32p, C.c_size_t,
                                               C.POINTER(C.c_size_t)]
    L.rsvio_tracker_remove_ids.argtypes = [H, C.c_int32, _u64p, C.c_size_t]
    L.rsvio_tracker_counters.argtypes = [H, _u64p]
    L.rsvio_tracker_overflow_frames.argtypes = [H, C.POINTER(C.c_uint64)]
    L.rsvio_tracker_set_profiling.argtypes = [H, C.c_int32]
    L.rsvio_tracker_stage_times.argtypes = [H, _f32p, _u64p, C.c_int32]
    L.rsvio_pyramid_bytes.restype = C.c_size_t
    L.rsvio_pyramid_bytes.argtypes = [C.c_int32] * 3
    L.rsvio_pyramid_device_bytes.restype = C.c_size_t
    L.rsvio_pyramid_device_bytes.argtypes = [C.c_int32] * 3
    L.rsvio_build_pyramid_u8.argtypes = [C.c_int32, _u8p, C.c_int32, C.c_int32, C.c_int32, C.c_size_t, C.c_size_t,
                                         C.c_int32, _u8p]
    L.rsvio_build_pyramid_u8_device.argtypes = [C.c_int32, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_size_t,
                                                C.c_size_t, C.c_int32, C.c_void_p, C.c_int32, C.POINTER(C.c_float)]
    L.rsvio_track_points.argtypes = [C.c_int32, _u8p, _u8p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, _f32p,
                                     C.c_int32, C.c_float, _f32p, _i32p, _i32p]
    L.rsvio_detect_keypoints.argtypes = [C.c_int32, _u8p, C.c_int32, C.c_int32, C.c_size_t, C.c_uint32, _u32p,
                                         C.c_int32, _u32p, _f32p, C.c_size_t, C.POINTER(C.c_size_t)]
    L.rsvio_se2_exp.argtypes = [C.c_int32, _f32p, C.c_int32, _f32p]
    L.rsvio_microbench_fp32.argtypes = [C.c_int32, C.POINTER(C.c_float), C.POINTER(C.c_float)]
    L.rsvio_tracker_set_camera.argtypes = [H, C.c_int32, C.POINTER(CameraModel)]
    L.rsvio_tracker_undistorted_view.argtypes = [H, C.POINTER(C.c_void_p), C.POINTER(C.c_int32), C.POINTER(C.c_int32)]
    L.rsvio_unproject_points.argtypes = [C.c_int32, C.POINTER(CameraModel), _f32p, C.c_int32, _f32p]
    L.rsvio_tracker_process_stereo_fmt.argtypes = [H, C.c_void_p, C.c_void_p, C.c_size_t, C.c_int32]
    L.rsvio_convert_to_luma8.argtypes = [C.c_int32, C.c_void_p, C.c_int32, C.c_int32, C.c_size_t, C.c_int32, _u8p]
    L.rsvio_b_pyramid_dims.argtypes = [C.c_int32, C.c_int32, C.c_int32, C.c_double, _i32p, _i32p]
    L.rsvio_b_build_pyramid.argtypes = [C.c_int32, _f32p, C.c_int32, C.c_int32, C.c_int32, C.c_double,
                                        C.POINTER(C.c_void_p)]
    L.rsvio_b_track_points.argtypes = [C.c_int32, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), _i32p, _i32p, C.c_int32,
                                       _f32p, C.c_int32, C.c_int32, C.c_float, _f32p, _i32p]
    L.rsvio_b_track_points_cost.argtypes = [C.c_int32, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), _i32p, _i32p, C.c_int32,
                                            _f32p, C.c_int32, C.c_int32, C.c_float, C.c_int32, _f32p, _i32p]
    L.rsvio_b_blur.argtypes = [C.c_int32, _f32p, C.c_int32, C.c_int32, C.c_float, _f32p]
    L.rsvio_b_shi_tomasi_score.argtypes = [C.c_int32, _f32p, C.c_int32, C.c_int32, C.c_float, _f32p]
    L.rsvio_b_add_points.argtypes = [C.c_int32, _f32p, C.c_int32, C.c_int32, _f32p, C.c_int32, C.c_float, C.c_uint32,
                                     C.c_float, np.ctypeslib.ndpointer(np.uint32, flags="C_CONTIGUOUS"), C.c_size_t,
                                     C.POINTER(C.c_size_t)]
    L.rsvio_b_config_default.argtypes = [C.POINTER(BConfig)]
    L.rsvio_b_config_default.restype = None
    L.rsvio_b_tracker_create.argtypes = [C.POINTER(BConfig), C.POINTER(C.c_void_p)]
    L.rsvio_b_tracker_destroy.argtypes = [C.c_void_p]
    L.rsvio_b_tracker_destroy.restype = None
    L.rsvio_b_tracker_process_frame.argtypes = [C.c_void_p, _f32p, C.c_int64]
    L.rsvio_b_tracker_num_features.argtypes = [C.c_void_p, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]
    L.rsvio_b_tracker_get_features.argtypes = [C.c_void_p, np.ctypeslib.ndpointer(np.uint64, flags="C_CONTIGUOUS"), _f32p,
                                               C.c_size_t, C.POINTER(C.c_size_t)]
    L.rsvio_b_tracker_pyramid_dims.argtypes = [C.c_void_p, C.POINTER(C.c_int32), _i32p, _i32p]
    L.rsvio_b_tracker_get_pyramid_level.argtypes = [C.c_void_p, C.c_int32, _f32p]
    L.rsvio_b_tracker_last_frame_ms.argtypes = [C.c_void_p, C.POINTER(C.c_float)]
    _lib = L
    return L


class CapacityWarning(RuntimeWarning):
    """new points did not fit the track table (rsvio_config.max_tracks); the frame itself was processed"""


def _check_frame(status, handle):
    """status of a process_* / sync call: a full track table is a warning, not an error -- the reference's maps
    are unbounded (feature_tracker.rs:93-96) and its process_frame cannot fail, so the frame is delivered."""
    if status == ERR_CAPACITY:
        import warnings
        warnings.warn(load_library().rsvio_tracker_last_error(handle).decode(), CapacityWarning, stacklevel=3)
        return
    _check(status, handle)


def _check(status, handle=None):
    if status != OK:
        L = load_library()
        msg = L.rsvio_status_string(status).decode()
        if handle:
            detail = L.rsvio_tracker_last_error(handle).decode()
            if detail:
                msg += ": " + detail
        raise RsvioError(status, msg)


def device_count():
    n = C.c_int32(0)
    load_library().rsvio_device_count(C.byref(n))
    return n.value


HOST_WRITE_COMBINED = 1


class HostBuffer:
    """Pinned host memory from rsvio_host_alloc[_flags] as a numpy uint8 array (`.array`), freed on close()."""

    def __init__(self, shape, write_combined=False):
        self.nbytes = int(np.prod(shape))
        self._p = C.c_void_p()
        _check(load_library().rsvio_host_alloc_flags(C.byref(self._p), self.nbytes,
                                                     HOST_WRITE_COMBINED if write_combined else 0))
        buf = (C.c_uint8 * self.nbytes).from_address(self._p.value)
        self.array = np.frombuffer(buf, dtype=np.uint8).reshape(shape)

    @property
    def ptr(self):
        return self._p.value

    def close(self):
        if self._p:
            self.array = None
            load_library().rsvio_host_free(self._p)
            self._p = C.c_void_p()


# ---- stateless entry points ---------------------------------------------------------------
def pyramid_bytes(w, h, levels):
    return load_library().rsvio_pyramid_bytes(w, h, levels)


def split_pyramid(buf, w0, h0, levels):
    out, off = [], 0
    for i in range(levels):
        w, h = w0 >> i, h0 >> i
        out.append(buf[off:off + w * h].reshape(h, w))
        off += w * h
    return out


def build_pyramid(images, levels, device=0):
    """images: (h, w) or (n, h, w) uint8.  Returns (n, pyramid_bytes) uint8 (flat dense pyramids)."""
    a = np.ascontiguousarray(images, np.uint8)
    if a.ndim == 2:
        a = a[None]
    n, h, w = a.shape
    out = np.empty((n, pyramid_bytes(w, h, levels)), np.uint8)
    _check(load_library().rsvio_build_pyramid_u8(device, a, n, w, h, w, w * h, levels, out))
    return out


def track_points(pyr_a, pyr_b, w0, h0, levels, in_T6, max_it, thr, device=0):
    in_T6 = np.ascontiguousarray(in_T6, np.float32).reshape(-1, 6)
    n = in_T6.shape[0]
    out_T = np.zeros((max(n, 1), 6), np.float32)
    status = np.zeros(max(n, 1), np.int32)
    iters = np.zeros((max(n, 1), 2), np.int32)
    _check(load_library().rsvio_track_points(device, np.ascontiguousarray(pyr_a, np.uint8),
                                             np.ascontiguousarray(pyr_b, np.uint8), w0, h0, levels, n, in_T6, max_it,
                                             thr, out_T, status, iters))
    return out_T[:n], status[:n], iters[:n]


def detect_keypoints(image, grid, cur_xy=None, device=0):
    img = np.ascontiguousarray(image, np.uint8)
    h, w = img.shape
    cur = np.zeros((0, 2), np.uint32) if cur_xy is None else np.ascontiguousarray(cur_xy, np.uint32).reshape(-1, 2)
    cap = max(1, (h // grid) * (w // grid))
    xy = np.zeros((cap, 2), np.uint32)
    sc = np.zeros(cap, np.float32)
    n = C.c_size_t(0)
    cur_arg = cur if cur.shape[0] else np.zeros((1, 2), np.uint32)
    _check(load_library().rsvio_detect_keypoints(device, img, w, h, w, grid, cur_arg, cur.shape[0], xy, sc, cap,
                                                 C.byref(n)))
    return xy[:n.value].copy(), sc[:n.value].copy()


def b_build_pyramid(image, levels, ratio=2.0, device=0):
    """crate B's build_image_pyramid (no blur) on the device: image (h, w) float32 in [0, 1] -> list of levels."""
    img = np.ascontiguousarray(image, np.float32)
    h, w = img.shape
    ws, hs = np.zeros(levels, np.int32), np.zeros(levels, np.int32)
    L = load_library()
    _check(L.rsvio_b_pyramid_dims(w, h, levels, float(ratio), ws, hs))
    out = [np.empty((int(hs[l]), int(ws[l])), np.float32) for l in range(levels)]
    ptrs = (C.c_void_p * levels)(*[a.ctypes.data for a in out])
    _check(L.rsvio_b_build_pyramid(device, img, w, h, levels, float(ratio), ptrs))
    return out


def b_track_points(pyr0, pyr1, points, max_iteration=10, lm_lambda=0.1, device=0):
    """crate B's track_points (row f-2, first slice): pyr0 / pyr1 = lists of (h, w) float32 levels, points (n, 2).
    Returns (T4[n, 4] = cos, sin, tx, ty of the forward isometry, status[n])."""
    lv0 = [np.ascontiguousarray(a, np.float32) for a in pyr0]
    lv1 = [np.ascontiguousarray(a, np.float32) for a in pyr1]
    L = len(lv0)
    if len(lv1) != L or any(a.shape != b.shape for a, b in zip(lv0, lv1)):
        raise ValueError("the two pyramids must have the same level shapes")
    p0 = (C.c_void_p * L)(*[a.ctypes.data for a in lv0])
    p1 = (C.c_void_p * L)(*[a.ctypes.data for a in lv1])
    ws = np.array([a.shape[1] for a in lv0], np.int32)
    hs = np.array([a.shape[0] for a in lv0], np.int32)
    pts = np.ascontiguousarray(points, np.float32).reshape(-1, 2)
    n = pts.shape[0]
    T = np.zeros((n, 4), np.float32)
    st = np.zeros(n, np.int32)
    _check(load_library().rsvio_b_track_points(device, p0, p1, ws, hs, L, pts, n, max_iteration, lm_lambda, T, st))
    return T, st


def b_track_points_cost(pyr0, pyr1, points, max_iteration=10, lm_lambda=0.1, cost=B_COST_SSD, device=0):
    """b_track_points with crate B's matching cost (B_COST_SSD | B_COST_LSSD, feature_tracker/src/patch.rs:4-8)"""
    lv0 = [np.ascontiguousarray(a, np.float32) for a in pyr0]
    lv1 = [np.ascontiguousarray(a, np.float32) for a in pyr1]
    L = len(lv0)
    if len(lv1) != L or any(a.shape != b.shape for a, b in zip(lv0, lv1)):
        raise ValueError("the two pyramids must have the same level shapes")
    p0 = (C.c_void_p * L)(*[a.ctypes.data for a in lv0])
    p1 = (C.c_void_p * L)(*[a.ctypes.data for a in lv1])
    ws = np.array([a.shape[1] for a in lv0], np.int32)
    hs = np.array([a.shape[0] for a in lv0], np.int32)
    pts = np.ascontiguousarray(points, np.float32).reshape(-1, 2)
    n = pts.shape[0]
    T = np.zeros((n, 4), np.float32)
    st = np.zeros(n, np.int32)
    _check(load_library().rsvio_b_track_points_cost(device, p0, p1, ws, hs, L, pts, n, max_iteration, lm_lambda, cost, T,
                                                    st))
    return T, st


def b_blur(image, sigma, device=0):
    """crate B's preprocessing blur (`image::imageops::blur`, image_operations.rs:58-62) on the device"""
    img = np.ascontiguousarray(image, np.float32)
    h, w = img.shape
    out = np.empty_like(img)
    _check(load_library().rsvio_b_blur(device, img, w, h, float(sigma), out))
    return out


def b_shi_tomasi_score(image, detection_blur, device=0):
    """crate B's shi_tomasi_score (feature_detection.rs:83-165) on the device"""
    img = np.ascontiguousarray(image, np.float32)
    h, w = img.shape
    out = np.empty_like(img)
    _check(load_library().rsvio_b_shi_tomasi_score(device, img, w, h, float(detection_blur), out))
    return out


def b_add_points(image, tracked_xy, threshold, min_dist, detection_blur, device=0):
    """crate B's add_points (feature_detection.rs:48-80): new corners (n, 2) uint32 in (y, x) order"""
    img = np.ascontiguousarray(image, np.float32)
    h, w = img.shape
    tr = np.ascontiguousarray(tracked_xy, np.float32).reshape(-1, 2)
    cap = ((w + 1) // 2) * ((h + 1) // 2)
    out = np.zeros((cap, 2), np.uint32)
    n = C.c_size_t(0)
    tr_arg = tr if tr.shape[0] else np.zeros((1, 2), np.float32)
    _check(load_library().rsvio_b_add_points(device, img, w, h, tr_arg, tr.shape[0], float(threshold), int(min_dist),
                                             float(detection_blur), out, cap, C.byref(n)))
    return out[:n.value].copy()


class FeatureTrackerB:
    """Mirror of crate B's `FeatureTracker` (feature_tracker/src/feature_tracker.rs:50-195) over rsvio_b_tracker_*:
    process_frame(image) -> [(feature_id, x, y)], get_pyramid() -> list of levels or None."""

    def __init__(self, width, height, nlevels=5, ratio=2.0, preprocessing_blur=True, preprocessing_blur_sigma=2.0,
                 detection_threshold=2.5, detection_min_dist=15, detection_blur=6.0, optical_flow_max_iter=25,
                 optical_flow_lm_lambda=0.1, matching_cost=B_COST_SSD, device=0, max_features=0):
        L = load_library()
        cfg = BConfig()
        L.rsvio_b_config_default(C.byref(cfg))
        cfg.width, cfg.height, cfg.nlevels, cfg.ratio = width, height, nlevels, ratio
        cfg.preprocessing_blur, cfg.preprocessing_blur_sigma = int(bool(preprocessing_blur)), preprocessing_blur_sigma
        cfg.detection_threshold, cfg.detection_min_dist = detection_threshold, detection_min_dist
        cfg.detection_blur, cfg.optical_flow_max_iter = detection_blur, optical_flow_max_iter
        cfg.optical_flow_lm_lambda, cfg.matching_cost, cfg.device = optical_flow_lm_lambda, matching_cost, device
        cfg.max_features = max_features
        self._L, self._h = L, C.c_void_p()
        self.width, self.height, self._frames = width, height, 0
        _check(L.rsvio_b_tracker_create(C.byref(cfg), C.byref(self._h)))

    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            self._L.rsvio_b_tracker_destroy(self._h)
            self._h = None

    __del__ = close

    def process_frame(self, image, frame_id=0):
        img = np.ascontiguousarray(image, np.float32)
        if img.shape != (self.height, self.width):
            raise ValueError("image shape does not match the tracker")
        _check(self._L.rsvio_b_tracker_process_frame(self._h, img, int(frame_id)))
        self._frames += 1
        return self.features()

    def features(self):
        n, n_new = C.c_int32(0), C.c_int32(0)
        _check(self._L.rsvio_b_tracker_num_features(self._h, C.byref(n), C.byref(n_new)))
        ids, xy = np.zeros(max(n.value, 1), np.uint64), np.zeros((max(n.value, 1), 2), np.float32)
        got = C.c_size_t(0)
        _check(self._L.rsvio_b_tracker_get_features(self._h, ids, xy, n.value, C.byref(got)))
        self.n_new = n_new.value
        return [(int(ids[i]), xy[i, 0], xy[i, 1]) for i in range(n.value)]

    def get_pyramid(self):
        if self._frames == 0:
            return None
        lv = C.c_int32(0)
        ws, hs = np.zeros(8, np.int32), np.zeros(8, np.int32)
        _check(self._L.rsvio_b_tracker_pyramid_dims(self._h, C.byref(lv), ws, hs))
        out = []
        for l in range(lv.value):
            a = np.empty((int(hs[l]), int(ws[l])), np.float32)
            _check(self._L.rsvio_b_tracker_get_pyramid_level(self._h, l, a))
            out.append(a)
        return out

    def last_frame_ms(self):
        ms = C.c_float(0)
        _check(self._L.rsvio_b_tracker_last_frame_ms(self._h, C.byref(ms)))
        return ms.value


def convert_to_luma8(img, fmt, device=0):
    """DynamicImage::to_luma8 on the device: (h, w) uint16 for PIXEL_LUMA16, (h, w, 3) uint8 for PIXEL_RGB8."""
    a = np.ascontiguousarray(img)
    h, w = a.shape[:2]
    out = np.empty((h, w), np.uint8)
    _check(load_library().rsvio_convert_to_luma8(device, C.c_void_p(a.ctypes.data), w, h, a.strides[0], fmt, out))
    return out


def unproject_points(model, xy, device=0):
    """camera.unproject_one for (n, 2) float32 pixels -> (n, 2) float32 normalised coordinates (NaN outside)."""
    xy = np.ascontiguousarray(xy, np.float32).reshape(-1, 2)
    out = np.empty_like(xy)
    _check(load_library().rsvio_unproject_points(device, C.byref(model), xy, xy.shape[0], out))
    return out


def microbench_fp32(device=0):
    """(FFMA TFLOP/s, separately rounded FMUL + FADD TFLOP/s) measured on the device right now"""
    a, b = C.c_float(0), C.c_float(0)
    _check(load_library().rsvio_microbench_fp32(device, C.byref(a), C.byref(b)))
    return a.value, b.value


def se2_exp(twists, device=0):
    tw = np.ascontiguousarray(twists, np.float32).reshape(-1, 3)
    out = np.zeros((tw.shape[0], 9), np.float32)
    _check(load_library().rsvio_se2_exp(device, tw, tw.shape[0], out))
    return out.reshape(-1, 3, 3)


# ---- stateful trackers ----------------------------------------------------------------------
@dataclass
class Feature:
    """reference Feature (feature_tracker.rs:12-32)"""
    feature_id: int
    pixel_coord: tuple
    undistorted_coord: tuple = field(default=(-1.0, -1.0))


class _Tracker:
    def __init__(self, width, height, levels, grid_size, max_iterations, convergence_threshold, mode,
                 n_streams=1, device=0, max_tracks=0, cuda_stream=None):
        L = load_library()
        cfg = Config()
        L.rsvio_config_default(C.byref(cfg))
        cfg.width, cfg.height, cfg.levels = width, height, levels
        cfg.grid_size, cfg.max_iterations = grid_size, max_iterations
        cfg.convergence_threshold = convergence_threshold
        cfg.mode, cfg.n_streams, cfg.device, cfg.max_tracks = mode, n_streams, device, max_tracks
        cfg.cuda_stream = cuda_stream
        self._h = C.c_void_p()
        self._L = L
        self.width, self.height, self.levels, self.n_streams, self.device = width, height, levels, n_streams, device
        self.stereo = mode == MODE_STEREO
        st = L.rsvio_tracker_create(C.byref(cfg), C.byref(self._h))
        if st != OK:
            self._h = None
            raise RsvioError(st, L.rsvio_status_string(st).decode() + " (rsvio_tracker_create)")
        cells = (width // grid_size) * (height // grid_size)
        self._cap = max_tracks if max_tracks > 0 else 4 * cells + 64
        self._keep = None

    def close(self):
        if getattr(self, "_h", None):
            self._L.rsvio_tracker_destroy(self._h)
            self._h = None

    def __del__(self):
        self.close()

    def reset(self):
        _check(self._L.rsvio_tracker_reset(self._h), self._h)

    @staticmethod
    def _ptrs(images, n, w, h):
        arrs = [np.ascontiguousarray(im, np.uint8) for im in images]
        if len(arrs) != n or any(a.shape != (h, w) for a in arrs):
            raise ValueError("expected %d images of shape (%d, %d)" % (n, h, w))
        return arrs, (C.c_void_p * n)(*[a.ctypes.data for a in arrs])

    def process_batch(self, left, right=None):
        la, lp = self._ptrs(left, self.n_streams, self.width, self.height)
        if self.stereo:
            ra, rp = self._ptrs(right, self.n_streams, self.width, self.height)
        else:
            ra, rp = None, None
        _check_frame(self._L.rsvio_tracker_process_batch(self._h, lp, rp, self.width), self._h)

    def submit_batch_ptr(self, left_ptrs, right_ptrs, stride=None):
        """left_ptrs/right_ptrs: ctypes arrays of host addresses (e.g. into pinned memory)."""
        _check(self._L.rsvio_tracker_submit_batch(self._h, left_ptrs, right_ptrs, stride or self.width), self._h)

    def submit_device(self, d_left, d_right, stride=None, pitch=None):
        stride = stride or self.width
        pitch = pitch or stride * self.height
        _check(self._L.rsvio_tracker_submit_device(self._h, C.c_void_p(d_left), C.c_void_p(d_right or 0), stride,
                                                   pitch), self._h)

    def sync(self):
        _check_frame(self._L.rsvio_tracker_sync(self._h), self._h)

    def overflow_frames(self):
        """frames since create/reset whose new points did not all fit the track table"""
        n = C.c_uint64(0)
        _check(self._L.rsvio_tracker_overflow_frames(self._h, C.byref(n)), self._h)
        return n.value

    def num_tracks(self, cam=0, stream=0):
        n = C.c_int32(0)
        _check(self._L.rsvio_tracker_num_tracks(self._h, stream, cam, C.byref(n)), self._h)
        return n.value

    def tracks(self, cam=0, stream=0):
        """(ids uint64[n], xy float32[n, 2]) of one camera, ids ascending."""
        ids = np.zeros(self._cap, np.uint64)
        xy = np.zeros((self._cap, 2), np.float32)
        n = C.c_size_t(0)
        _check(self._L.rsvio_tracker_get_tracks(self._h, stream, cam, ids, xy, self._cap, C.byref(n)), self._h)
        return ids[:n.value].copy(), xy[:n.value].copy()

    RECORD_DTYPE = np.dtype([("id", np.uint64), ("x", np.float32), ("y", np.float32)])

    def result_view(self):
        """Zero-copy numpy views of the last synced frame: (records[groups, cap], counts[groups]) living
        in the host-mapped pinned ring the merge kernel writes to; group = cam * n_streams + stream."""
        rec, cnt = C.c_void_p(), C.c_void_p()
        groups, cap = C.c_int32(0), C.c_int32(0)
        _check(self._L.rsvio_tracker_result_view(self._h, C.byref(rec), C.byref(cnt), C.byref(groups), C.byref(cap)),
               self._h)
        g, c = groups.value, cap.value
        rbuf = (C.c_char * (g * c * 16)).from_address(rec.value)
        cbuf = (C.c_int32 * g).from_address(cnt.value)
        return (np.frombuffer(rbuf, dtype=self.RECORD_DTYPE).reshape(g, c), np.frombuffer(cbuf, dtype=np.int32))

    def set_camera(self, cam, model):
        """camera model of cam 0 (left) / 1 (right): from the next frame on every record also carries its
        unprojected coordinate (Feature.undistorted_coord, estimator/frame.rs:107-134)."""
        _check(self._L.rsvio_tracker_set_camera(self._h, cam, C.byref(model)), self._h)

    def undistorted_view(self):
        """uxy[groups, cap, 2] float32 (zero-copy) for the records of result_view(), or None without a model."""
        ptr = C.c_void_p()
        groups, cap = C.c_int32(0), C.c_int32(0)
        _check(self._L.rsvio_tracker_undistorted_view(self._h, C.byref(ptr), C.byref(groups), C.byref(cap)), self._h)
        if not ptr.value:
            return None
        g, c = groups.value, cap.value
        buf = (C.c_float * (g * c * 2)).from_address(ptr.value)
        return np.frombuffer(buf, dtype=np.float32).reshape(g, c, 2)

    def transforms(self, cam=0, stream=0):
        ids = np.zeros(self._cap, np.uint64)
        T = np.zeros((self._cap, 6), np.float32)
        n = C.c_size_t(0)
        _check(self._L.rsvio_tracker_get_transforms(self._h, stream, cam, ids, T, self._cap, C.byref(n)), self._h)
        return ids[:n.value].copy(), T[:n.value].copy()

    def remove_id(self, ids, stream=0):
        ids = np.ascontiguousarray(ids, np.uint64)
        if ids.size:
            _check(self._L.rsvio_tracker_remove_ids(self._h, stream, ids, ids.size), self._h)

    def counters(self):
        out = np.zeros(6, np.uint64)
        _check(self._L.rsvio_tracker_counters(self._h, out), self._h)
        return dict(tracks=int(out[0]), frames=int(out[1]), iterations=int(out[2]), templates=int(out[3]),
                    launches=int(out[4]), detected=int(out[5]))

    def set_profiling(self, on=True):
        _check(self._L.rsvio_tracker_set_profiling(self._h, 1 if on else 0), self._h)

    def stage_times(self, reset=True):
        ms = np.zeros(N_STAGES, np.float32)
        ln = np.zeros(N_STAGES, np.uint64)
        _check(self._L.rsvio_tracker_stage_times(self._h, ms, ln, 1 if reset else 0), self._h)
        return {k: (float(ms[i]), int(ln[i])) for i, k in enumerate(STAGE_NAMES)}


class StereoPatchTracker(_Tracker):
    """StereoPatchTracker<LEVELS> (feature_tracker.rs:91-207) on the GPU."""

    def __init__(self, grid_size, optical_flow_max_iterations, optical_flow_convergence_threshold, *,
                 width, height, levels=6, n_streams=1, device=0, max_tracks=0, cuda_stream=None):
        super().__init__(width, height, levels, grid_size, optical_flow_max_iterations,
                         float(optical_flow_convergence_threshold), MODE_STEREO, n_streams, device, max_tracks,
                         cuda_stream)

    def process_frame(self, greyscale_image0, greyscale_image1, frame=None):
        """feature_tracker.rs:116-187.  `frame` (optional) needs add_left_feature / add_right_feature."""
        l = np.ascontiguousarray(greyscale_image0, np.uint8)
        r = np.ascontiguousarray(greyscale_image1, np.uint8)
        if l.shape != (self.height, self.width) or r.shape != l.shape:
            raise ValueError("image shape does not match the tracker")
        _check_frame(self._L.rsvio_tracker_process_stereo(self._h, C.c_void_p(l.ctypes.data),
                                                          C.c_void_p(r.ctypes.data), self.width), self._h)
        if frame is not None:
            uxy = self.undistorted_view()
            for cam, add in ((0, frame.add_left_feature), (1, frame.add_right_feature)):
                ids, xy = self.tracks(cam)
                und = uxy[cam * self.n_streams, :len(ids)].tolist() if uxy is not None else None
                for k, (i, p) in enumerate(zip(ids.tolist(), xy.tolist())):
                    # with a camera model set the unprojection of Frame::add_*_feature (frame.rs:118-119) is
                    # already done on the device; otherwise the reference's placeholder (-1, -1) stays
                    add(Feature(i, (p[0], p[1]), tuple(und[k]) if und is not None else (-1.0, -1.0)))

    def process_decoded(self, left, right, fmt):
        """process_frame for frames as the decoder delivers them: (h, w) uint16 (PIXEL_LUMA16) or (h, w, 3) uint8
        (PIXEL_RGB8); the to_luma8() of the reference's dataset players runs on the device."""
        l, r = np.ascontiguousarray(left), np.ascontiguousarray(right)
        if l.shape[:2] != (self.height, self.width) or r.shape != l.shape:
            raise ValueError("image shape does not match the tracker")
        _check_frame(self._L.rsvio_tracker_process_stereo_fmt(self._h, C.c_void_p(l.ctypes.data),
                                                              C.c_void_p(r.ctypes.data), l.strides[0], fmt), self._h)

    def get_track_points(self, stream=0):
        """[ {id: (x, y)} for cam0, cam1 ]  (feature_tracker.rs:188-200)"""
        out = []
        for cam in (0, 1):
            ids, xy = self.tracks(cam, stream)
            out.append({int(i): (float(p[0]), float(p[1])) for i, p in zip(ids, xy)})
        return out


class PatchTracker(_Tracker):
    """PatchTracker<LEVELS> (feature_tracker.rs:36-89); defaults 30 iterations / 0.005 / grid 30."""

    def __init__(self, *, width, height, levels=6, grid_size=30, max_iterations=30, convergence_threshold=0.005,
                 n_streams=1, device=0, max_tracks=0, cuda_stream=None):
        super().__init__(width, height, levels, grid_size, max_iterations, convergence_threshold, MODE_MONO,
                         n_streams, device, max_tracks, cuda_stream)

    def process_frame(self, greyscale_image):
        im = np.ascontiguousarray(greyscale_image, np.uint8)
        if im.shape != (self.height, self.width):
            raise ValueError("image shape does not match the tracker")
        _check_frame(self._L.rsvio_tracker_process_mono(self._h, C.c_void_p(im.ctypes.data), self.width), self._h)

    def get_track_points(self, stream=0):
        ids, xy = self.tracks(0, stream)
        return {int(i): (float(p[0]), float(p[1])) for i, p in zip(ids, xy)}
