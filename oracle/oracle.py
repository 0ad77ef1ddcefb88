"""ctypes binding of oracle/lib/librsvio_oracle.so (TEST INFRASTRUCTURE ONLY).

The C++ oracle restates RS-VIO tracker A (reference src/feature_tracker/*.rs); see
rsvio_oracle.h for the per-function file:line citations.  PARITY UNPINNED.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "lib", "librsvio_oracle.so")
NP = 52

TRACK_OK, FAIL_TEMPLATE, FAIL_RESIDUAL, FAIL_NONFINITE, FAIL_DIVERGED, FAIL_OOB, FAIL_FB = range(7)
FAIL_BACKWARD = 16


def build(force=False):
    """Compile the oracle with its Makefile (g++, -ffp-contract=off)."""
    src = [os.path.join(_HERE, f) for f in ("rsvio_oracle.cpp", "rsvio_oracle.h", "Makefile")]
    if not force and os.path.exists(_LIB_PATH) and all(
            os.path.getmtime(_LIB_PATH) >= os.path.getmtime(s) for s in src):
        return _LIB_PATH
    subprocess.run(["make", "-C", _HERE], check=True, stdout=subprocess.DEVNULL)
    return _LIB_PATH


class Pattern52(C.Structure):
    _fields_ = [("valid", C.c_int32), ("mean", C.c_float), ("data", C.c_float * NP),
                ("hinv_jt", C.c_float * (3 * NP))]


_u8p = np.ctypeslib.ndpointer(np.uint8, flags="C_CONTIGUOUS")
_f32p = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")
_i32p = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")
_u32p = np.ctypeslib.ndpointer(np.uint32, flags="C_CONTIGUOUS")
_u64p = np.ctypeslib.ndpointer(np.uint64, flags="C_CONTIGUOUS")

_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_LIB_PATH):
        build()
    L = C.CDLL(_LIB_PATH)
    L.orc_pyramid_bytes.restype = C.c_size_t
    L.orc_pyramid_bytes.argtypes = [C.c_int] * 3
    L.orc_resize_triangle_u8.argtypes = [_u8p, C.c_int, C.c_int, _u8p, C.c_int, C.c_int]
    L.orc_build_pyramid.argtypes = [_u8p, C.c_int, C.c_int, C.c_int, _u8p, C.c_int]
    L.orc_se2_exp.argtypes = [_f32p, _f32p]
    L.orc_image_grad.argtypes = [_u8p, C.c_int, C.c_int, C.c_float, C.c_float, _f32p]
    L.orc_inbound.argtypes = [C.c_int, C.c_int, C.c_float, C.c_float, C.c_uint32]
    L.orc_pattern52_new.argtypes = [_u8p, C.c_int, C.c_int, C.c_float, C.c_float, C.POINTER(Pattern52)]
    L.orc_pattern52_residual.argtypes = [C.POINTER(Pattern52), _u8p, C.c_int, C.c_int, _f32p, _f32p]
    L.orc_track_point_at_level.argtypes = [_u8p, C.c_int, C.c_int, C.POINTER(Pattern52), _f32p, C.c_int,
                                           C.c_float, C.POINTER(C.c_int)]
    L.orc_track_points.argtypes = [_u8p, _u8p, C.c_int, C.c_int, C.c_int, C.c_int, _f32p, C.c_int, C.c_float,
                                   _f32p, _i32p, _i32p, _f32p, C.c_int]
    L.orc_corners_fast9.argtypes = [_u8p, C.c_int, C.c_int, C.c_int, _u32p, _u32p, _f32p, C.c_int]
    L.orc_detect_key_points.argtypes = [_u8p, C.c_int, C.c_int, C.c_uint32, _u32p, C.c_int, _u32p, _u32p, _f32p,
                                        C.c_int]
    L.orc_tracker_create.restype = C.c_void_p
    L.orc_tracker_create.argtypes = [C.c_int, C.c_uint32, C.c_int, C.c_float, C.c_int]
    L.orc_tracker_destroy.argtypes = [C.c_void_p]
    L.orc_tracker_process_stereo.argtypes = [C.c_void_p, _u8p, _u8p, C.c_int, C.c_int]
    L.orc_tracker_process_mono.argtypes = [C.c_void_p, _u8p, C.c_int, C.c_int]
    L.orc_tracker_num_tracks.argtypes = [C.c_void_p, C.c_int]
    L.orc_tracker_get_tracks.argtypes = [C.c_void_p, C.c_int, _u64p, _f32p, C.c_int]
    L.orc_tracker_remove_ids.argtypes = [C.c_void_p, _u64p, C.c_int]
    L.orc_tracker_counters.argtypes = [C.c_void_p, _u64p]
    _lib = L
    return L


def level_dims(w0, h0, levels):
    return [(w0 >> i, h0 >> i) for i in range(levels)]


def pyramid_bytes(w0, h0, levels):
    return sum(w * h for w, h in level_dims(w0, h0, levels))


def split_pyramid(buf, w0, h0, levels):
    out, off = [], 0
    for w, h in level_dims(w0, h0, levels):
        out.append(buf[off:off + w * h].reshape(h, w))
        off += w * h
    return out


def resize_triangle(img, nw, nh):
    img = np.ascontiguousarray(img, np.uint8)
    h, w = img.shape
    out = np.empty((nh, nw), np.uint8)
    lib().orc_resize_triangle_u8(img, w, h, out, nw, nh)
    return out


def build_pyramid(img, levels, nthreads=1):
    """Returns the flat pyramid buffer (levels back to back, dense row-major)."""
    img = np.ascontiguousarray(img, np.uint8)
    h, w = img.shape
    out = np.empty(pyramid_bytes(w, h, levels), np.uint8)
    lib().orc_build_pyramid(img, w, h, levels, out, nthreads)
    return out


def se2_exp(a):
    a = np.ascontiguousarray(a, np.float32)
    out = np.empty(9, np.float32)
    lib().orc_se2_exp(a, out)
    return out.reshape(3, 3)


PIXEL_LUMA8, PIXEL_LUMA16, PIXEL_RGB8 = 0, 1, 2


def to_luma8(img, fmt):
    """DynamicImage::to_luma8: img = (h, w) uint16 for PIXEL_LUMA16, (h, w, 3) uint8 for PIXEL_RGB8."""
    a = np.ascontiguousarray(img)
    h, w = a.shape[:2]
    out = np.empty((h, w), np.uint8)
    f = lib().orc_to_luma8
    f.restype = None
    f.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_size_t, C.c_int, _u8p]
    f(a.ctypes.data, w, h, a.strides[0], fmt, out)
    return out


CAMERA_PINHOLE_RADTAN, CAMERA_EUCM = 1, 2


def unproject(model, params, uv):
    """camera.unproject_one for an (n, 2) array of pixels -> (n, 2) float64 normalised coordinates (NaN outside)."""
    P = (C.c_double * 9)(*(list(params) + [0.0] * 9)[:9])
    uv = np.asarray(uv, np.float64).reshape(-1, 2)
    out = np.full(uv.shape, np.nan)
    o = (C.c_double * 2)()
    f = lib().orc_unproject
    f.restype = C.c_int
    for i, (u, v) in enumerate(uv):
        if f(C.c_int(model), P, C.c_double(u), C.c_double(v), o):
            out[i] = (o[0], o[1])
    return out


def project(model, params, xy):
    P = (C.c_double * 9)(*(list(params) + [0.0] * 9)[:9])
    xy = np.asarray(xy, np.float64).reshape(-1, 2)
    out = np.full(xy.shape, np.nan)
    o = (C.c_double * 2)()
    f = lib().orc_project
    f.restype = C.c_int
    for i, (x, y) in enumerate(xy):
        if f(C.c_int(model), P, C.c_double(x), C.c_double(y), o):
            out[i] = (o[0], o[1])
    return out


def image_grad(img, x, y):
    img = np.ascontiguousarray(img, np.uint8)
    out = np.empty(3, np.float32)
    lib().orc_image_grad(img, img.shape[1], img.shape[0], x, y, out)
    return out


def inbound(w, h, x, y, radius):
    return bool(lib().orc_inbound(w, h, x, y, radius))


def pattern52_new(img, px, py):
    img = np.ascontiguousarray(img, np.uint8)
    p = Pattern52()
    lib().orc_pattern52_new(img, img.shape[1], img.shape[0], px, py, C.byref(p))
    return p


def pattern52_residual(pat, img, xy):
    img = np.ascontiguousarray(img, np.uint8)
    xy = np.ascontiguousarray(xy, np.float32).reshape(2, NP)
    res = np.empty(NP, np.float32)
    ok = lib().orc_pattern52_residual(C.byref(pat), img, img.shape[1], img.shape[0], xy, res)
    return (res if ok else None)


def track_point_at_level(img, pat, T, max_it, thr):
    img = np.ascontiguousarray(img, np.uint8)
    T = np.array(T, np.float32).reshape(6).copy()
    it = C.c_int(0)
    st = lib().orc_track_point_at_level(img, img.shape[1], img.shape[0], C.byref(pat), T, max_it, thr, C.byref(it))
    return st, T, it.value


def track_points(pyr_a, pyr_b, w0, h0, levels, in_T, max_it, thr, nthreads=1, want_margins=True):
    in_T = np.ascontiguousarray(in_T, np.float32).reshape(-1, 6)
    n = in_T.shape[0]
    out_T = np.zeros((n, 6), np.float32)
    status = np.zeros(n, np.int32)
    iters = np.zeros((n, 2), np.int32)
    margins = np.zeros((n, 3), np.float32)
    lib().orc_track_points(pyr_a, pyr_b, w0, h0, levels, n, in_T, max_it, thr, out_T, status, iters, margins,
                           nthreads)
    return out_T, status, iters, margins


def corners_fast9(img, threshold):
    img = np.ascontiguousarray(img, np.uint8)
    cap = img.size
    xs, ys, sc = np.empty(cap, np.uint32), np.empty(cap, np.uint32), np.empty(cap, np.float32)
    n = lib().orc_corners_fast9(img, img.shape[1], img.shape[0], threshold, xs, ys, sc, cap)
    return xs[:n].copy(), ys[:n].copy(), sc[:n].copy()


def detect_key_points(img, grid, cur_xy):
    img = np.ascontiguousarray(img, np.uint8)
    cur = np.ascontiguousarray(cur_xy, np.uint32).reshape(-1, 2)
    cap = (img.shape[0] // grid + 1) * (img.shape[1] // grid + 1)
    xs, ys, sc = np.empty(cap, np.uint32), np.empty(cap, np.uint32), np.empty(cap, np.float32)
    n = lib().orc_detect_key_points(img, img.shape[1], img.shape[0], grid, cur, cur.shape[0], xs, ys, sc, cap)
    return xs[:n].copy(), ys[:n].copy(), sc[:n].copy()


def identity_transforms(xy):
    xy = np.asarray(xy, np.float32).reshape(-1, 2)
    T = np.zeros((xy.shape[0], 6), np.float32)
    T[:, 0] = 1.0
    T[:, 4] = 1.0
    T[:, 2] = xy[:, 0]
    T[:, 5] = xy[:, 1]
    return T


class StereoTracker:
    """Oracle StereoPatchTracker<L> (feature_tracker.rs:91-207), ids in ascending detection order."""

    def __init__(self, levels, grid, max_it, thr, nthreads=1):
        self._h = lib().orc_tracker_create(levels, grid, max_it, thr, nthreads)

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_tracker_destroy(self._h)
            self._h = None

    def process_frame(self, left, right):
        left = np.ascontiguousarray(left, np.uint8)
        right = np.ascontiguousarray(right, np.uint8)
        lib().orc_tracker_process_stereo(self._h, left, right, left.shape[1], left.shape[0])

    def process_mono(self, img):
        img = np.ascontiguousarray(img, np.uint8)
        lib().orc_tracker_process_mono(self._h, img, img.shape[1], img.shape[0])

    def get_tracks(self, cam):
        n = lib().orc_tracker_num_tracks(self._h, cam)
        ids = np.empty(max(n, 1), np.uint64)
        T = np.empty((max(n, 1), 6), np.float32)
        lib().orc_tracker_get_tracks(self._h, cam, ids, T, n)
        return ids[:n], T[:n]

    def remove_ids(self, ids):
        ids = np.ascontiguousarray(ids, np.uint64)
        lib().orc_tracker_remove_ids(self._h, ids, ids.size)

    def counters(self):
        out = np.zeros(2, np.uint64)
        lib().orc_tracker_counters(self._h, out)
        return int(out[0]), int(out[1])
