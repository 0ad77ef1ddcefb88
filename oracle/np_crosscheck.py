"""Second, independently written restatement (numpy, vectorised) of the two pieces of the hot
path whose arithmetic lives in un-vendored crates, used to pin the C++ oracle byte for byte.

TEST INFRASTRUCTURE ONLY.  PARITY UNPINNED (no reference golden vectors exist).

* resize_triangle  -- image 0.25 `imageops::resize(.., FilterType::Triangle)` on Luma<u8>
  (call sites: reference src/feature_tracker/feature_tracker.rs:210,217)
* fast9_score_map  -- imageproc 0.25 `corners_fast9` + `fast_corner_score`, in closed form
  (call site: reference src/feature_tracker/image_utilities.rs:156)
* detect_key_points -- reference src/feature_tracker/image_utilities.rs:108-175
* lk_level          -- Pattern52::new + track_point_at_level in plain Python/np.float32
  (reference src/feature_tracker/patch.rs:75-232, feature_tracker.rs:344-395)
"""
import math

import numpy as np

F = np.float32


def _taps(n_in, n_out):
    """Per output index: first source index, and the normalised f32 tent weights."""
    ratio = F(n_in) / F(n_out)
    sratio = ratio if ratio >= F(1) else F(1)
    support = F(1.0) * sratio
    out = []
    for o in range(n_out):
        c = (F(o) + F(0.5)) * ratio
        left = int(math.floor(c - support))
        left = min(max(left, 0), n_in - 1)
        right = int(math.ceil(c + support))
        right = min(max(right, left + 1), n_in)
        c = c - F(0.5)
        idx = np.arange(left, right).astype(F)
        x = np.abs((idx - c) / sratio).astype(F)
        w = np.where(x < F(1), F(1) - x, F(0)).astype(F)
        s = F(0)
        for v in w:  # sequential f32 sum
            s = F(s + v)
        out.append((left, (w / s).astype(F)))
    return out


def resize_triangle(img, nw, nh):
    img = np.asarray(img, np.uint8)
    h, w = img.shape
    if (nw, nh) == (w, h):
        return img.copy()
    src = img.astype(F)
    tmp = np.zeros((nh, w), F)
    for oy, (top, wy) in enumerate(_taps(h, nh)):
        acc = np.zeros(w, F)
        for k, wk in enumerate(wy):  # tap order, multiply then add, all f32
            acc = (acc + (src[top + k] * wk).astype(F)).astype(F)
        tmp[oy] = acc
    out = np.zeros((nh, nw), np.uint8)
    for ox, (left, wx) in enumerate(_taps(w, nw)):
        acc = np.zeros(nh, F)
        for k, wk in enumerate(wx):
            acc = (acc + (tmp[:, left + k] * wk).astype(F)).astype(F)
        acc = np.clip(acc, F(0), F(255))
        # round half away from zero; acc >= 0 so floor(acc + 0.5) computed in f64 is exact
        out[:, ox] = np.floor(acc.astype(np.float64) + 0.5).astype(np.uint8)
    return out


def build_pyramid(img, levels):
    h, w = img.shape
    return [resize_triangle(img, w >> i, h >> i) for i in range(levels)]


_CIRCLE = [(0, -3), (1, -3), (2, -2), (3, -1), (3, 0), (3, 1), (2, 2), (1, 3),
           (0, 3), (-1, 3), (-2, 2), (-3, 1), (-3, 0), (-3, -1), (-2, -2), (-1, -3)]


def fast9_score_map(img):
    """score[y, x] = largest t in [0, 254] at which (x, y) is a FAST-9 corner, -1 if none
    (closed form of the binary search: max over the 16 arcs of 9 of the arc's min contrast, minus 1).
    A pixel is reported by corners_fast9(img, thr) iff score >= thr, with that score."""
    img = np.asarray(img, np.uint8).astype(np.int32)
    h, w = img.shape
    score = np.full((h, w), -1, np.int32)
    if h < 7 or w < 7:
        return score
    c = img[3:h - 3, 3:w - 3]
    ring = np.stack([img[3 + dy:h - 3 + dy, 3 + dx:w - 3 + dx] for dx, dy in _CIRCLE])  # 16 x H x W
    bright = ring - c[None]
    dark = c[None] - ring
    best = np.full(c.shape, -1, np.int32)
    for start in range(16):
        idx = [(start + k) % 16 for k in range(9)]
        best = np.maximum(best, bright[idx].min(axis=0) - 1)
        best = np.maximum(best, dark[idx].min(axis=0) - 1)
    score[3:h - 3, 3:w - 3] = best
    return score


def corners_fast9(img, threshold):
    s = fast9_score_map(img)
    ys, xs = np.nonzero(s >= threshold)  # raster order (y outer, x inner)
    return xs.astype(np.uint32), ys.astype(np.uint32), s[ys, xs].astype(np.float32)


def detect_key_points(img, grid, cur_xy):
    img = np.asarray(img, np.uint8)
    h, w = img.shape
    g = int(grid)
    edge = 19
    occ = np.zeros((h // g + 1, w // g + 1), np.int32)
    x_start = (w % g) // 2
    x_stop = x_start + g * (w // g - 1) + 1
    y_start = (h % g) // 2
    y_stop = y_start + g * (h // g - 1) + 1
    for cx, cy in np.asarray(cur_xy, np.int64).reshape(-1, 2):
        if x_start <= cx < x_stop + g and y_start <= cy < y_stop + g:
            occ[(cy - y_start) // g, (cx - x_start) // g] += 1
    out = []
    for x in range(x_start, x_stop, g):
        for y in range(y_start, y_stop, g):
            if occ[(y - y_start) // g, (x - x_start) // g] > 0:
                continue
            s = fast9_score_map(img[y:y + g, x:x + g])
            ys, xs = np.nonzero(s >= 10)
            sc = s[ys, xs]
            gx, gy = xs + x, ys + y
            ok = (gx >= edge) & (gx <= w - edge) & (gy >= edge) & (gy <= h - edge)
            for thr in range(40, 5, -5):
                m = ok & (sc >= thr)
                if m.any():
                    k = np.flatnonzero(m)
                    j = k[np.argmin(sc[k], axis=0)]  # first minimum in raster order == stable sort
                    out.append((int(gx[j]), int(gy[j]), float(sc[j])))
                    break
    return out


# ---------------------------------------------------------------------------------------------
# One LK level in scalar np.float32 (slow; for a handful of points)
# ---------------------------------------------------------------------------------------------
def pattern_raw():
    half_w = [3, 5, 7, 7, 7, 7, 5, 3]
    pts = []
    for r, hw in enumerate(half_w):
        y = 7 - 2 * r
        for x in range(-hw, hw + 1, 2):
            pts.append((x, y))
    return np.array(pts, F)


def _bilin(img, ix, iy, dx, dy):
    ddx, ddy = F(1) - dx, F(1) - dy
    p = lambda a, b: F(img[iy + b, ix + a])
    return F(F(F(F(ddx * ddy) * p(0, 0)) + F(F(ddx * dy) * p(0, 1))) + F(F(dx * ddy) * p(1, 0))) + F(F(dx * dy) * p(1, 1))


def _round_u32(v):
    r = math.floor(abs(float(v)) + 0.5) * (1 if v >= 0 else -1)
    return int(max(r, 0))


def _inbound(w, h, x, y, r):
    rx, ry = _round_u32(x), _round_u32(y)
    return rx >= r and ry >= r and rx < w - r and ry < h - r


def pattern52_new(img, px, py):
    h, w = img.shape
    raw = pattern_raw()
    px, py = F(px), F(py)
    data = np.zeros(52, F)
    J = np.zeros((52, 3), F)
    s = F(0)
    gs = np.zeros(3, F)
    n = 0
    for i in range(52):
        ox, oy = F(raw[i, 0] / F(2)), F(raw[i, 1] / F(2))
        x, y = F(px + ox), F(py + oy)
        if _inbound(w, h, x, y, 2):
            ix, iy = int(math.floor(x)), int(math.floor(y))
            dx, dy = F(x - F(ix)), F(y - F(iy))
            v = F(_bilin(img, ix, iy, dx, dy))
            gx = F(F(0.5) * F(F(_bilin(img, ix + 1, iy, dx, dy)) - F(_bilin(img, ix - 1, iy, dx, dy))))
            gy = F(F(0.5) * F(F(_bilin(img, ix, iy + 1, dx, dy)) - F(_bilin(img, ix, iy - 1, dx, dy))))
            data[i] = v
            s = F(s + v)
            J[i] = (gx, gy, F(F(gx * F(-oy)) + F(gy * ox)))
            gs = (gs + J[i]).astype(F)
            n += 1
        else:
            data[i] = F(-1)
    with np.errstate(all="ignore"):
        mean = F(s / F(n)) if n else F(np.nan)
        mean_inv = F(F(n) / s)
        for i in range(52):
            if data[i] >= 0:
                rhs = ((gs * data[i]).astype(F) / s).astype(F)
                J[i] = (J[i] - rhs).astype(F)
                data[i] = F(data[i] * mean_inv)
            else:
                J[i] = 0
        J = (J * mean_inv).astype(F)
        H = np.zeros((3, 3), F)
        for a in range(3):
            for b in range(3):
                acc = F(J[0, a] * J[0, b])
                for i in range(1, 52):
                    acc = F(acc + F(J[i, a] * J[i, b]))
                H[a, b] = acc
    return dict(data=data, J=J, H=H, mean=mean, n=n)
