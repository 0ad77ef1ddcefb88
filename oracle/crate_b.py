"""CPU restatement (numpy, f32) of the reference's SECOND tracker, the `feature_tracker` crate ("crate B",
SURVEY.md 8 row f-2).  TEST INFRASTRUCTURE ONLY: the product never imports it.

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
  Patch52 (LSSD)               patch.rs:84-105, 167-216
  shi_tomasi_score             feature_tracker/src/feature_tracker/feature_detection.rs:83-165
  suppress_non_maximum         feature_detection.rs:172-254 (in-tree copy of imageproc's block NMS + threshold)
  add_points                   feature_detection.rs:48-80
  FeatureTracker               feature_tracker/src/feature_tracker.rs:50-195
NOT pinned (arithmetic of crates that are not vendored; restated from their published sources, PARITY UNPINNED):
  `image` 0.25.9  imageops::resize on Luma<f32> (as Appendix A of SURVEY.md: no rounding, clamp to [0, 1] -- inputs
                  are expected in [0, 1], which is what `to_luma32f` delivers, players/tartanair_player.rs:52);
                  imageops::blur restated as the sampling implementation (Gaussian kernel, support 2 sigma, vertical then
                  horizontal pass, weights normalised by their sum) -- later 0.25.x releases may convolve with a fixed
                  kernel instead; imageops::fast_blur restated as three box passes (sizes from `boxes_for_gauss`), each a
                  horizontal running sum written transposed, edges replicated, every value clamped to the sample type's
                  [0, 1] range (so the blurred Ix*Iy plane is >= 0);
  `imageproc` 0.26  filter::horizontal_filter / vertical_filter with [-1, 0, 1] (correlation, edges replicated);
                  suppress::local_maxima (sorted by (y, x); rows [cy - r, min(cy + r + 1, last_row)) are searched, i.e.
                  the last occupied row is never searched; ties resolved lexicographically)."""
import ctypes
import ctypes.util
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
def _expf(x):
    """f32::exp of the reference = the platform libm's expf (numpy's own float32 exp may differ by an ulp)"""
    global _LIBM
    try:
        _LIBM
    except NameError:
        _LIBM = ctypes.CDLL(ctypes.util.find_library("m") or "libm.so.6")
        _LIBM.expf.restype = ctypes.c_float
        _LIBM.expf.argtypes = [ctypes.c_float]
    return F(_LIBM.expf(float(x)))


def _tent(x):
    xx = abs(x)
    return F(1.0) - xx if xx < F(1.0) else F(0.0)


def _sample_axis_f32(img, n_out, axis, kernel=_tent, support=F(1.0)):
    """one pass of image::imageops::{vertical,horizontal}_sample on an f32 image along `axis`,
    no rounding"""
    img = np.moveaxis(np.asarray(img, np.float32), axis, 0)
    n_in = img.shape[0]
    ratio = F(n_in) / F(n_out)
    sratio = max(ratio, F(1.0))
    src_support = F(support) * sratio
    out = np.zeros((n_out,) + img.shape[1:], np.float32)
    for o in range(n_out):
        c = (F(o) + F(0.5)) * ratio
        left = min(max(int(math.floor(c - src_support)), 0), n_in - 1)
        right = min(max(int(math.ceil(c + src_support)), left + 1), n_in)
        cc = c - F(0.5)
        ws = [kernel((F(i) - cc) / sratio) for i in range(left, right)]
        s = F(0.0)
        for wv in ws:
            s = s + wv
        t = np.zeros(img.shape[1:], np.float32)
        for k, i in enumerate(range(left, right)):
            t = t + img[i] * (ws[k] / s)
        out[o] = t
    return np.moveaxis(out, 0, axis)


def _resize_axis_f32(img, n_out, axis):
    return _sample_axis_f32(img, n_out, axis)


def gaussian_kernel(sigma):
    """imageops/sample.rs `gaussian(x, r)`: ((2 pi).sqrt() * r).recip() * (-x.powi(2) / (2.0 * r.powi(2))).exp()"""
    r = F(sigma)
    norm = F(1.0) / (F(np.sqrt(F(2.0) * F(math.pi))) * r)
    den = F(2.0) * (r * r)
    return lambda x: norm * _expf(-(F(x) * F(x)) / den)


def blur(img, sigma):
    """image::imageops::blur (PARITY UNPINNED, see the module docstring): Filter{gaussian, support 2 sigma}, vertical
    pass into un-clamped f32, horizontal pass, clamp to [0, 1]"""
    sigma = F(1.0) if sigma <= 0 else F(sigma)
    h, w = np.asarray(img).shape
    k = gaussian_kernel(sigma)
    tmp = _sample_axis_f32(img, h, 0, k, F(2.0) * sigma)
    return np.clip(_sample_axis_f32(tmp, w, 1, k, F(2.0) * sigma), F(0.0), F(1.0)).astype(np.float32)


def resize_triangle_f32(img, nw, nh):
    """vertical pass first, then horizontal, then the Luma<f32> clamp to [0, 1]"""
    tmp = _resize_axis_f32(img, nh, 0)
    return np.clip(_resize_axis_f32(tmp, nw, 1), F(0.0), F(1.0)).astype(np.float32)


def _round_half_away(v):
    return int(math.floor(abs(v) + 0.5)) * (1 if v >= 0 else -1)


def build_image_pyramid(img, nlevels, ratio=2.0, bluring=False, blur_sigma=0.0):
    """image_operations.rs:47-78"""
    pyr = [blur(img, blur_sigma) if bluring else np.asarray(img, np.float32)]
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


def track_point_at_level(img0, img1, coords, T0, max_iteration, lm_lambda, cost="ssd"):
    """feature_tracking.rs:129-192; returns the transform or None (OutOfBound)"""
    T = Iso2(cs=(T0.c, T0.s), t=T0.t.copy())
    patch = (Patch52Lssd if cost == "lssd" else Patch52)(img0, coords, lm_lambda)
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


def track_point(pyr0, pyr1, point, max_iteration, lm_lambda, cost="ssd"):
    """feature_tracking.rs:69-124"""
    h, w = pyr0[0].shape
    T = Iso2()
    for level in range(len(pyr0) - 1, -1, -1):
        hl, wl = pyr0[level].shape
        sx, sy = F(wl) / F(w), F(hl) / F(h)
        coords = (sx * (F(point[0]) + F(0.5)) - F(0.5), sy * (F(point[1]) + F(0.5)) - F(0.5))
        T = track_point_at_level(pyr0[level], pyr1[level], coords, T, max_iteration, lm_lambda, cost)
        if T is None:
            return None
        if level > 0:
            hn, wn = pyr0[level - 1].shape
            T.t[0] = T.t[0] * (F(wn) / F(wl))
            T.t[1] = T.t[1] * (F(hn) / F(hl))
    return T


def track_points(pyr0, pyr1, points, max_iteration=10, lm_lambda=0.1, cost="ssd"):
    """feature_tracking.rs:16-59: {index: transform} of the points that pass the forward-backward test (< 2 px)"""
    out = {}
    for k, p in enumerate(points):
        T0 = track_point(pyr0, pyr1, p, max_iteration, lm_lambda, cost)
        if T0 is None:
            continue
        p1 = T0.apply(p)
        T1 = track_point(pyr1, pyr0, p1, max_iteration, lm_lambda, cost)
        if T1 is None:
            continue
        back = T1.apply(p1)
        if F(np.linalg.norm(np.array(p, np.float32) - back)) < F(2.0):
            out[k] = T0
    return out


# ---- LSSD cost (patch.rs:84-105, 167-216; public through Patch52::new / residuals, process_frame hard-codes SSD) ----
class Patch52Lssd(Patch52):
    """compute_intensities_and_jacobian_lssd + residuals_lssd.  As in the reference, `intensities` stay RAW (the
    template is not divided by its mean) while the residual divides the target samples by THEIR mean."""

    def __init__(self, img, center, lm_lambda):
        self.center = np.array(center, np.float32)
        self.intensities = np.zeros(52, np.float32)
        dI = np.zeros((52, 2), np.float32)
        mean_i, mean_d = F(0.0), np.zeros(2, np.float32)
        g = np.zeros(2, np.float32)
        for i, sh in enumerate(PATTERN_RAW):
            px, py = self.center[0] + sh[0], self.center[1] + sh[1]
            r = d_interpolate_bicubic(img, px, py)
            if r is None:
                self.intensities[i] = F(0.0)          # unwrap_or(0.0); dimg_dpixel keeps its last value
            else:
                self.intensities[i], g = r
            dI[i] = g
            mean_i = mean_i + self.intensities[i]
            mean_d = (mean_d + g).astype(np.float32)
        mean_i = mean_i / F(52.0)
        mean_d = (mean_d / F(52.0)).astype(np.float32)
        self.J = np.zeros((52, 3), np.float32)
        m2 = mean_i * mean_i
        for i, sh in enumerate(PATTERN_RAW):
            px, py = self.center[0] + sh[0], self.center[1] + sh[1]
            d = ((dI[i] * mean_i - self.intensities[i] * mean_d) / m2).astype(np.float32)
            self.J[i] = (d[0] * (-py) + d[1] * px, d[0], d[1])
        H = np.eye(3, dtype=np.float32) * F(lm_lambda)
        H = (H + self.J.T @ self.J).astype(np.float32)
        self.Hinv = inverse3_f32(H)

    def residuals(self, img, T):
        r = np.zeros(52, np.float32)
        for i, sh in enumerate(PATTERN_RAW):
            p = T.apply((self.center[0] + sh[0], self.center[1] + sh[1]))
            v = interpolate_bicubic(img, p[0], p[1])
            r[i] = F(0.0) if v is None else v
        s = F(0.0)
        for v in r:                                   # nalgebra mean(): sequential sum / len
            s = s + v
        r = (r / (s / F(52.0))).astype(np.float32)
        return (r - self.intensities).astype(np.float32)


# ---- detection (feature_detection.rs) ------------------------------------------------------------------------------
def boxes_for_gauss(sigma, n=3):
    """image imageops/fast_blur.rs: box sizes whose n-fold convolution approximates a Gaussian of std sigma (f32)"""
    sigma = F(sigma)
    w_ideal = F(np.sqrt(F(12.0) * (sigma * sigma) / F(n) + F(1.0)))
    w_l = F(np.floor(w_ideal))
    if w_l % F(2.0) == 0:
        w_l = w_l - F(1.0)
    w_u = w_l + F(2.0)
    m_ideal = F(0.25) * F(n) * (w_l + F(3.0)) - F(3.0) * (sigma * sigma) * (F(1.0) / (w_l + F(1.0)))
    m = _round_half_away(float(m_ideal))
    m = max(m, 0)
    return [int(w_l) if i < m else int(w_u) for i in range(n)]


def _box_pass_transposed(a, r):
    """horizontal_fast_blur_half: running sums along the rows of `a` (edges replicated), every value divided by
    2r + 1 and clamped to [0, 1], result written TRANSPOSED.  Vectorised over the rows; the sum order along a row is
    the reference's."""
    h, w = a.shape
    out = np.zeros((w, h), np.float32)
    cl = lambda x: min(max(x, 0), w - 1)
    vals = np.zeros(h, np.float32)
    for x in range(-r, r + 1):
        vals = vals + a[:, cl(x)]
    den = F(2.0) * F(r) + F(1.0)
    for c in range(w):
        out[c, :] = np.clip(vals / den, F(0.0), F(1.0))
        vals = vals - a[:, cl(c - r)] + a[:, cl(c + r + 1)]
    return out


def fast_blur(img, sigma):
    """image::imageops::fast_blur (PARITY UNPINNED): 3 box passes, each horizontal then (on the transposed data)
    horizontal again"""
    a = np.asarray(img, np.float32)
    if a.size == 0:
        return a.copy()
    for size in boxes_for_gauss(sigma, 3):
        r = (size - 1) // 2
        a = _box_pass_transposed(_box_pass_transposed(a, r), r)
    return a


def shi_tomasi_score(img, detection_blur):
    """feature_detection.rs:83-165"""
    a = np.asarray(img, np.float32)
    p = np.pad(a, 1, mode="edge")
    dx = (p[1:-1, 2:] - p[1:-1, :-2]).astype(np.float32)      # horizontal_filter [-1, 0, 1]: -p(x-1) + 0 + p(x+1)
    dy = (p[2:, 1:-1] - p[:-2, 1:-1]).astype(np.float32)
    dxx = fast_blur(dx * dx, detection_blur)
    dyy = fast_blur(dy * dy, detection_blur)
    dxy = fast_blur(dx * dy, detection_blur)
    trace = dxx + dyy
    det = dxx * dyy - dxy * dxy
    delta = np.maximum(trace * trace - F(4.0) * det, F(0.0)).astype(np.float32)
    return (F(500.0) * np.abs(trace - np.sqrt(delta))).astype(np.float32)


def suppress_non_maximum(score, radius, threshold):
    """feature_detection.rs:172-254: [(x, y, score)] in block order (y outer, x inner)"""
    h, w = score.shape
    out = []
    step = radius + 1

    def greater(x, y, v, y_lo, y_hi, x_lo, x_hi):
        for cy in range(y_lo, y_hi):
            for cx in range(x_lo, x_hi):
                ci = score[cy, cx]
                if ci < v:
                    continue
                if ci > v or (cx, cy) < (x, y):
                    return True
        return False

    for y in range(0, h, step):
        for x in range(0, w, step):
            bx, by, bs = x, y, score[y, x]
            for cy in range(y, min(h, y + radius + 1)):
                for cx in range(x, min(w, x + radius + 1)):
                    ci = score[cy, cx]
                    if ci < bs:
                        continue
                    if ci > bs or (cx, cy) < (bx, by):
                        bx, by, bs = cx, cy, ci
            if bs >= threshold:
                x0, x1, x2, x3 = max(bx - radius, 0), x, min(w, x + radius + 1), min(w, bx + radius + 1)
                y0, y1, y2, y3 = max(by - radius, 0), y, min(h, y + radius + 1), min(h, by + radius + 1)
                failed = greater(bx, by, bs, y0, y1, x0, x3)
                failed |= greater(bx, by, bs, y1, y2, x0, x1)
                failed |= greater(bx, by, bs, y1, y2, x2, x3)
                failed |= greater(bx, by, bs, y2, y3, x0, x3)
                if not failed:
                    out.append((bx, by, F(bs)))
    return out


def suppress_non_maximum_r1(score, threshold):
    """the same for radius 1, vectorised (used on full-size images); identical output incl. order"""
    h, w = score.shape
    out = []
    S = np.asarray(score, np.float32)
    for y in range(0, h, 2):
        for x in range(0, w, 2):
            blk = S[y:y + 2, x:x + 2]
            # best: maximal score, ties to the lexicographically smallest (cx, cy) -- the reference compares (x, y) tuples
            m = blk.max()
            cand = [(cx, cy) for cy in range(blk.shape[0]) for cx in range(blk.shape[1]) if blk[cy, cx] == m]
            bx, by = min(cand)
            bx, by = bx + x, by + y
            if not m >= threshold:
                continue
            y0, y3, x0, x3 = max(by - 1, 0), min(h, by + 2), max(bx - 1, 0), min(w, bx + 2)
            win = S[y0:y3, x0:x3]
            if (win > m).any():
                continue
            eq = np.argwhere(win == m)
            if any(((cx + x0, cy + y0) < (bx, by)) and not (y <= cy + y0 < min(h, y + 2) and x <= cx + x0 < min(w, x + 2))
                   for cy, cx in eq):
                continue
            out.append((bx, by, F(m)))
    return out


def local_maxima(items, radius):
    """imageproc::suppress::local_maxima (PARITY UNPINNED).  items: [(x, y, score, tag)]; returns the kept ones in
    (y, x) order (stable)."""
    ordered = sorted(items, key=lambda c: (c[1], c[0]))
    if not ordered:
        return []
    height = ordered[-1][1]
    by_row = [[] for _ in range(height + 1)]
    for t in ordered:
        by_row[t[1]].append(t)
    keep = []
    for t in ordered:
        cx, cy, cs = t[0], t[1], t[2]
        is_max = True
        row_lower = 0 if radius > cy else cy - radius
        row_upper = height if cy + radius + 1 > height else cy + radius + 1
        for y in range(row_lower, row_upper):
            for c in by_row[y]:
                if c[0] + radius < cx:
                    continue
                if c[0] > cx + radius:
                    break
                if c[2] > cs:
                    is_max = False
                    break
                if c[2] < cs:
                    continue
                if (c[1], c[0]) < (cy, cx):
                    is_max = False
                    break
            if not is_max:
                break
        if is_max:
            keep.append(t)
    return keep


def add_points(level0, tracked_xy, threshold, min_dist, detection_blur):
    """feature_detection.rs:48-80: new corners [(x, y)] (u32) in (y, x) order"""
    score = shi_tomasi_score(level0, detection_blur)
    corners = [(x, y, s, False) for x, y, s in suppress_non_maximum_r1(score, F(threshold))]
    if len(tracked_xy) > 0:
        mx = F(-np.inf)
        for c in corners:
            mx = max(mx, c[2])
        for p in tracked_xy:
            corners.append((max(_round_half_away(float(p[0])), 0), max(_round_half_away(float(p[1])), 0), mx + F(1.0), True))
    kept = local_maxima(corners, int(min_dist))
    h, w = np.asarray(level0).shape
    return [(c[0], c[1]) for c in kept
            if not c[3] and min_dist <= c[0] < w - min_dist and min_dist <= c[1] < h - min_dist]


class FeatureTracker:
    """feature_tracker.rs:50-195 (viewer hooks dropped).  features: list of (id, x, y)."""

    def __init__(self, nlevels=5, ratio=2.0, preprocessing_blur=True, preprocessing_blur_sigma=2.0,
                 detection_threshold=2.5, detection_min_dist=15, detection_blur=6.0, optical_flow_max_iter=25,
                 optical_flow_lm_lambda=0.1, matching_cost="ssd"):
        self.cost = matching_cost   # process_frame hard-codes SSD (feature_tracker.rs:125); "lssd" mirrors the handle's option
        self.cfg = dict(nlevels=nlevels, ratio=ratio, blur=preprocessing_blur, sigma=preprocessing_blur_sigma,
                        thr=detection_threshold, min_dist=detection_min_dist, dblur=detection_blur,
                        max_iter=optical_flow_max_iter, lam=optical_flow_lm_lambda)
        self.prev = None            # (features, pyramid)
        self.last_id = 0

    def process_frame(self, image):
        c = self.cfg
        pyr = build_image_pyramid(image, c["nlevels"], c["ratio"], c["blur"], c["sigma"])
        feats = []
        if self.prev is not None:
            pf, pp = self.prev
            pts = [(x, y) for _, x, y in pf]
            maps = track_points(pp, pyr, pts, c["max_iter"], c["lam"], self.cost)
            for k, (fid, x, y) in enumerate(pf):
                if k in maps:
                    q = maps[k].apply((x, y))
                    feats.append((fid, F(q[0]), F(q[1])))
        new = add_points(pyr[0], [(x, y) for _, x, y in feats], c["thr"], c["min_dist"], c["dblur"])
        for x, y in new:
            feats.append((self.last_id, F(x), F(y)))
            self.last_id += 1
        self.prev = (feats, pyr)
        return feats

    def get_pyramid(self):
        return None if self.prev is None else self.prev[1]
