"""CPU restatement (numpy, f32) of the tracking core of the reference's SECOND tracker, the `feature_tracker`
crate ("crate B").  TEST INFRASTRUCTURE ONLY -- groundwork for row f-2, no CUDA path uses it
yet and the product never imports it.

Follows, function by function (paths relative to the reference repository):
  exp_se2 / log_se2            feature_tracker/src/feature_tracker/feature_tracking.rs:195-243   twist = [theta, vx, vy]
  bicubic_1d & derivatives     feature_tracker/src/image_operations.rs:232-282
  interpolate_bicubic          image_operations.rs:140-176     valid iff floor(x) in [1, w-3] and floor(y) in [1, h-3]
  d_interpolate_bicubic        image_operations.rs:181-229
  build_image_pyramid          image_operations.rs:47-78       level l = resize(level l-1, round(w0 / ratio^l), ..., Triangle)
  Patch52 (SSD)                feature_tracker/src/patch.rs:119-164, 244-255, 259-285   offsets NOT halved (radius 7 px)
  track_point_at_level         feature_tracking.rs:129-192
  track_point / track_points   feature_tracking.rs:16-124      forward-backward distance < 2 px
Pinned by the reference's own unit tests: the exp_se2 known answers and exp/log round trips
(feature_tracking.rs:264-297), the finite-difference property of the bicubic derivatives
(image_operations.rs:305-368) and the pyramid sizes (image_operations.rs:84-94) -- tests/test_crate_b_oracle.py.
NOT pinned: `image::imageops::resize` on f32 images (image 0.25.9, not vendored; restated as in Appendix A of
SURVEY.md: no rounding, clamp to [0, 1] -- so inputs are expected in [0, 1]) and `imageops::blur` (not restated
here: preprocessing_blur is not supported by this oracle yet).  Detection (Shi-Tomasi, fast_blur, NMS) is not
restated either."""
import math

import numpy as np

F = np.float32

def _pattern():
    """the 52 offsets of patch.rs:259-277: rows y = 7, 5, .., -7 with 4, 6, 8, 8, 8, 8, 6, 4 samples at odd x, left to
    right (raw_indexes() returns them un-halved: the patch spans +-7 px)"""
    half_widths = (3, 5, 7, 7, 7, 7, 5, 3)
    return np.array([[x, 7 - 2 * r] for r, hw in enumerate(half_widths) for x in range(-hw, hw + 1, 2)], np.float32)


PATTERN_RAW = _pattern()


# ---- SE(2) --------------------------------------------------------------------------------------------------
class Iso2:
    """na::Isometry2<f32>: rotation angle (unit complex) + translation; acts as R p + t."""

    def __init__(self, theta=0.0, t=(0.0, 0.0), cs=None):
        if cs is None:
            cs = (F(math.cos(F(theta))), F(math.sin(F(theta))))
        self.c, self.s = F(cs[0]), F(cs[1])
        self.t = np.array(t, np.float32)

    def apply(self, p):
        x, y = F(p[0]), F(p[1])
        return np.array([self.c * x - self.s * y + self.t[0], self.s * x + self.c * y + self.t[1]], np.float32)

    def mul(self, o):
        """self * o (apply o first)"""
        c = self.c * o.c - self.s * o.s
        s = self.s * o.c + self.c * o.s
        t = self.apply(o.t)
        return Iso2(cs=(c, s), t=t)

    def matrix(self):
        return np.array([[self.c, -self.s, self.t[0]], [self.s, self.c, self.t[1]], [0, 0, 1]], np.float32)


def exp_se2(twist):
    """feature_tracking.rs:195-219"""
    theta, v0, v1 = F(twist[0]), F(twist[1]), F(twist[2])
    if abs(theta) > F(1e-4):
        diag = F(math.sin(theta)) / theta
        cross = (F(1.0) - F(math.cos(theta))) / theta
    else:
        theta_sq = theta * theta
        diag = F(1.0) - theta * theta / F(6.0)
        cross = (F(0.5) - theta_sq / F(24.0)) * theta
    return Iso2(theta, (diag * v0 - cross * v1, cross * v0 + diag * v1))


def log_se2(T):
    """feature_tracking.rs:221-243"""
    theta = F(math.atan2(T.s, T.c))
    if abs(theta) > F(1e-3):
        diag = theta * F(math.sin(theta)) / (F(2.0) * (F(1.0) - F(math.cos(theta))))
    else:
        theta_sq = theta * theta
        diag = (F(1.0) - theta_sq / F(6.0)) / (F(1.0) - theta_sq / F(12.0))
    cross = F(0.5) * theta
    return np.array([theta, diag * T.t[0] + cross * T.t[1], -cross * T.t[0] + diag * T.t[1]], np.float32)


# ---- bicubic sampler (Catmull-Rom) ------------------------------------------------------------------------------
def _coeffs(f):
    a0 = f[1]
    a1 = f[2] - f[0]
    a2 = F(2.0) * f[0] - F(5.0) * f[1] + F(4.0) * f[2] - f[3]
    a3 = F(3.0) * (f[1] - f[2]) + f[3] - f[0]
    return a0, a1, a2, a3


def bicubic_1d(f, t):
    f = np.asarray(f, np.float32)
    t = F(t)
    a0, a1, a2, a3 = _coeffs(f)
    return a0 + F(0.5) * (t * (a1 + t * (a2 + t * a3)))


def d_bicubic_1d(f, t):
    """(value, d/dt)   image_operations.rs:246-261"""
    f = np.asarray(f, np.float32)
    t = F(t)
    a0, a1, a2, a3 = _coeffs(f)
    val = a0 + F(0.5) * (t * (a1 + t * (a2 + t * a3)))
    return val, F(0.5) * (a1 + t * (F(2.0) * a2 + t * F(3.0) * a3))


def d_bicubic_1d_full(f, t):
    """(value, d/df[4], d/dt)   image_operations.rs:263-282"""
    val, dt = d_bicubic_1d(f, t)
    t = F(t)
    df = np.array([F(0.5) * (t * (F(-1.0) + t * (F(2.0) + t * F(-1.0)))),
                   F(1.0) + F(0.5) * (t * (t * (F(-5.0) + t * F(3.0)))),
                   F(0.5) * (t * (F(1.0) + t * (F(4.0) + t * F(-3.0)))),
                   F(0.5) * (t * (t * (F(-1.0) + t)))], np.float32)
    return val, df, dt


def _window(img, x, y):
    h, w = img.shape
    x, y = F(x), F(y)
    if not (np.isfinite(x) and np.isfinite(y)):
        return None
    xf, yf = math.floor(x), math.floor(y)
    xf_u, yf_u = max(xf, 0), max(yf, 0)              # `as u32` saturates negatives to 0
    if not (1 <= xf_u <= max(w - 3, 0)) or not (1 <= yf_u <= max(h - 3, 0)):
        return None
    return xf_u, yf_u, x - F(xf_u), y - F(yf_u)


def interpolate_bicubic(img, x, y):
    """image_operations.rs:140-176; None when out of bounds"""
    wdw = _window(img, x, y)
    if wdw is None:
        return None
    xf, yf, tx, ty = wdw
    fy = [bicubic_1d(img[yf - 1 + dy, xf - 1:xf + 3], tx) for dy in range(4)]
    return bicubic_1d(fy, ty)


def d_interpolate_bicubic(img, x, y):
    """(value, [d/dx, d/dy]) or None   image_operations.rs:181-229"""
    wdw = _window(img, x, y)
    if wdw is None:
        return None
    xf, yf, tx, ty = wdw
    fy, dfy = [], []
    for dy in range(4):
        v, d = d_bicubic_1d(img[yf - 1 + dy, xf - 1:xf + 3], tx)
        fy.append(v)
        dfy.append(d)
    out, dout_dfy, dout_dty = d_bicubic_1d_full(fy, ty)
    dout_dtx = dout_dfy[0] * dfy[0] + dout_dfy[1] * dfy[1] + dout_dfy[2] * dfy[2] + dout_dfy[3] * dfy[3]
    return out, np.array([dout_dtx, dout_dty], np.float32)


def in_bounds(img, x, y):
    """image_operations.rs:4-7"""
    h, w = img.shape
    x, y = F(x), F(y)
    if not (x > 0 and y > 0):                         # is_positive(): +0.0 counts as positive, NaN does not
        if not ((x == 0 and not np.signbit(x)) or x > 0) or not ((y == 0 and not np.signbit(y)) or y > 0):
            return False
    rx, ry = int(np.floor(abs(x) + F(0.5))), int(np.floor(abs(y) + F(0.5)))   # round half away (values >= 0 here)
    return rx < w and ry < h


# ---- pyramid ----------------------------------------------------------------------------------------------------
def _resize_axis_f32(img, n_out, axis):
    """one pass of image::imageops::resize(Triangle) on an f32 image along `axis`, no rounding"""
    img = np.moveaxis(np.asarray(img, np.float32), axis, 0)
    n_in = img.shape[0]
    ratio = F(n_in) / F(n_out)
    sratio = max(ratio, F(1.0))
    support = F(1.0) * sratio
    out = np.zeros((n_out,) + img.shape[1:], np.float32)
    for o in range(n_out):
        c = (F(o) + F(0.5)) * ratio
        left = min(max(int(math.floor(c - support)), 0), n_in - 1)
        right = min(max(int(math.ceil(c + support)), left + 1), n_in)
        cc = c - F(0.5)
        ws = []
        for i in range(left, right):
            xx = abs((F(i) - cc) / sratio)
            ws.append(F(1.0) - xx if xx < F(1.0) else F(0.0))
        s = F(0.0)
        for wv in ws:
            s = s + wv
        t = np.zeros(img.shape[1:], np.float32)
        for k, i in enumerate(range(left, right)):
            t = t + img[i] * (ws[k] / s)
        out[o] = t
    return np.moveaxis(out, 0, axis)


def resize_triangle_f32(img, nw, nh):
    """vertical pass first, then horizontal, then the Luma<f32> clamp to [0, 1]"""
    tmp = _resize_axis_f32(img, nh, 0)
    return np.clip(_resize_axis_f32(tmp, nw, 1), F(0.0), F(1.0)).astype(np.float32)


def _round_half_away(v):
    return int(math.floor(abs(v) + 0.5)) * (1 if v >= 0 else -1)


def build_image_pyramid(img, nlevels, ratio=2.0):
    """image_operations.rs:47-78 without the preprocessing blur"""
    pyr = [np.asarray(img, np.float32)]
    h0, w0 = pyr[0].shape
    for l in range(1, nlevels):
        nw, nh = _round_half_away(w0 / ratio ** l), _round_half_away(h0 / ratio ** l)
        pyr.append(resize_triangle_f32(pyr[-1], nw, nh))
    return pyr


def inverse3_f32(m):
    """nalgebra `try_inverse_mut` of a 3x3 matrix: cofactors over the determinant, all in f32"""
    m = np.asarray(m, np.float32)
    m11, m12, m13 = m[0]
    m21, m22, m23 = m[1]
    m31, m32, m33 = m[2]
    minor_m12_m23 = m22 * m33 - m32 * m23
    minor_m11_m23 = m21 * m33 - m31 * m23
    minor_m11_m22 = m21 * m32 - m31 * m22
    det = m11 * minor_m12_m23 - m12 * minor_m11_m23 + m13 * minor_m11_m22
    if det == 0:
        raise ZeroDivisionError("Hessian should be invertible by its definition")   # the reference panics
    inv = np.array([
        [minor_m12_m23 / det, (m13 * m32 - m33 * m12) / det, (m12 * m23 - m22 * m13) / det],
        [-minor_m11_m23 / det, (m11 * m33 - m31 * m13) / det, (m13 * m21 - m23 * m11) / det],
        [minor_m11_m22 / det, (m12 * m31 - m32 * m11) / det, (m11 * m22 - m21 * m12) / det]], np.float32)
    return inv


# ---- Patch52 (SSD) and the tracker -----------------------------------------------------------------------------------
class Patch52:
    """patch.rs:119-164, 244-255"""

    def __init__(self, img, center, lm_lambda):
        self.center = np.array(center, np.float32)
        self.intensities = np.zeros(52, np.float32)
        self.J = np.zeros((52, 3), np.float32)
        for i, sh in enumerate(PATTERN_RAW):
            px, py = self.center[0] + sh[0], self.center[1] + sh[1]
            r = d_interpolate_bicubic(img, px, py)
            if r is None:                             # unwrap_or(0.0): value 0, gradient left over from the last sample
                self.intensities[i] = F(0.0)          # (the reference reuses dimg_dpixel; reproduced)
                g = self._last_g if hasattr(self, "_last_g") else np.zeros(2, np.float32)
            else:
                self.intensities[i], g = r
                self._last_g = g
            # [gx gy] * [[-py, 1, 0], [px, 0, 1]]
            self.J[i] = (g[0] * (-py) + g[1] * px, g[0], g[1])
        H = np.eye(3, dtype=np.float32) * F(lm_lambda)
        H = (H + self.J.T @ self.J).astype(np.float32)
        self.Hinv = inverse3_f32(H)

    def residuals(self, img, T):
        r = np.zeros(52, np.float32)
        for i, sh in enumerate(PATTERN_RAW):
            p = T.apply((self.center[0] + sh[0], self.center[1] + sh[1]))
            v = interpolate_bicubic(img, p[0], p[1])
            r[i] = (F(0.0) if v is None else v) - self.intensities[i]
        return r


def track_point_at_level(img0, img1, coords, T0, max_iteration, lm_lambda):
    """feature_tracking.rs:129-192; returns the transform or None (OutOfBound)"""
    T = Iso2(cs=(T0.c, T0.s), t=T0.t.copy())
    patch = Patch52(img0, coords, lm_lambda)
    for _ in range(max_iteration):
        res = patch.residuals(img1, T)
        b = (patch.J.T @ res).astype(np.float32)
        twist = (patch.Hinv @ b).astype(np.float32)
        T = T.mul(exp_se2(-twist))
        it = T.apply(coords)
        if not in_bounds(img0, it[0], it[1]):
            return None
        if F(np.linalg.norm(twist)) < F(1e-3):
            break
    return T


def track_point(pyr0, pyr1, point, max_iteration, lm_lambda):
    """feature_tracking.rs:69-124"""
    h, w = pyr0[0].shape
    T = Iso2()
    for level in range(len(pyr0) - 1, -1, -1):
        hl, wl = pyr0[level].shape
        sx, sy = F(wl) / F(w), F(hl) / F(h)
        coords = (sx * (F(point[0]) + F(0.5)) - F(0.5), sy * (F(point[1]) + F(0.5)) - F(0.5))
        T = track_point_at_level(pyr0[level], pyr1[level], coords, T, max_iteration, lm_lambda)
        if T is None:
            return None
        if level > 0:
            hn, wn = pyr0[level - 1].shape
            T.t[0] = T.t[0] * (F(wn) / F(wl))
            T.t[1] = T.t[1] * (F(hn) / F(hl))
    return T


def track_points(pyr0, pyr1, points, max_iteration=10, lm_lambda=0.1):
    """feature_tracking.rs:16-59: {index: transform} of the points that pass the forward-backward test (< 2 px)"""
    out = {}
    for k, p in enumerate(points):
        T0 = track_point(pyr0, pyr1, p, max_iteration, lm_lambda)
        if T0 is None:
            continue
        p1 = T0.apply(p)
        T1 = track_point(pyr1, pyr0, p1, max_iteration, lm_lambda)
        if T1 is None:
            continue
        back = T1.apply(p1)
        if F(np.linalg.norm(np.array(p, np.float32) - back)) < F(2.0):
            out[k] = T0
    return out
