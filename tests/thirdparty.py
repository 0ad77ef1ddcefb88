"""Third-party pins for arithmetic the reference delegates to crates that are not in its tree.

The Rust reference cannot be built in this image (no cargo), but the image carries INDEPENDENT implementations of
the same published algorithms.  Nothing in here shares code with oracle/rsvio_oracle.cpp or oracle/np_crosscheck.py:

* OpenCV `FastFeatureDetector(TYPE_9_16, nonmaxSuppression=False)` -- the same strict-inequality 9-of-16 segment
  test as imageproc::corners::corners_fast9 (reference call site src/feature_tracker/image_utilities.rs:156).
  A pixel's score "largest threshold at which it still is a corner" follows from the corner SETS alone
  (score(p) = max{t : p in FAST9(t)}), so OpenCV also pins every score and, through them, which corner
  detect_key_points picks in a cell (image_utilities.rs:151-170).
* OpenCV `undistortPoints` -- inverse of the radial-tangential model of row f-1.
* Pillow `resize(BILINEAR)` (antialiased tent of the same support and centre convention as image 0.25's
  Triangle filter, fixed-point coefficients, horizontal pass first) and a plain integer tent sum.
"""
import numpy as np


def cv_fast9_sets(img, t_lo=10, t_hi=255):
    """{t: boolean corner map of OpenCV FAST-9/16 at threshold t (no NMS)} for t_lo <= t <= t_hi, stopping at the
    first empty set (corner sets are nested: a corner at t is a corner at every smaller threshold)."""
    import cv2
    out = {}
    for t in range(t_lo, t_hi + 1):
        det = cv2.FastFeatureDetector_create(threshold=t, nonmaxSuppression=False,
                                             type=cv2.FAST_FEATURE_DETECTOR_TYPE_9_16)
        kps = det.detect(img, None)
        m = np.zeros(img.shape, bool)
        if kps:
            pts = np.array([k.pt for k in kps])
            m[pts[:, 1].astype(int), pts[:, 0].astype(int)] = True
        out[t] = m
        if not kps:
            break
    return out


def cv_fast9_score_map(img, t_lo=10):
    """score(p) = largest t >= t_lo at which OpenCV calls p a FAST-9 corner, -1 where it never does."""
    score = np.full(img.shape, -1, np.int32)
    for t, m in cv_fast9_sets(img, t_lo).items():
        score[m] = t
    return score


def cv_detect_key_points(img, grid, occupied_xy):
    """detect_key_points (image_utilities.rs:108-175, one point per cell) driven ONLY by OpenCV's corner sets:
    occupancy grid, cells x-outer / y-inner, thresholds 40, 35, .., 10, candidates whose FAST circle fits the
    g x g cell copy, stable ascending sort by score (raster order breaks ties), first one inside the 19-px border."""
    h, w = img.shape
    g = int(grid)
    score = cv_fast9_score_map(img, 10)
    x0, y0 = (w % g) // 2, (h % g) // 2
    occ = np.zeros((h // g + 1, w // g + 1), bool)
    for x, y in np.asarray(occupied_xy, np.int64).reshape(-1, 2):
        # image_utilities.rs:121-137: x_stop = x_start + g (w/g - 1) + 1, test x < x_stop + g
        if x >= x0 and y >= y0 and x < x0 + g * (w // g - 1) + 1 + g and y < y0 + g * (h // g - 1) + 1 + g:
            occ[(y - y0) // g, (x - x0) // g] = True
    out = []
    for cx in range(w // g):
        for cy in range(h // g):
            if occ[cy, cx]:
                continue
            ox, oy = x0 + cx * g, y0 + cy * g
            cell = score[oy + 3:oy + g - 3, ox + 3:ox + g - 3]
            found = None
            for t in range(40, 5, -5):
                ys, xs = np.nonzero(cell >= t)               # raster order (y outer)
                if len(xs) == 0:
                    continue
                order = np.argsort(cell[ys, xs], kind="stable")
                for k in order:
                    px, py = ox + 3 + int(xs[k]), oy + 3 + int(ys[k])
                    if 19 <= px <= w - 19 and 19 <= py <= h - 19:
                        found = (px, py, float(cell[ys[k], xs[k]]))
                        break
                if found:
                    break
            if found:
                out.append(found)
    return out


def tent_level_int(img, level):
    """Interior of pyramid level `level` (1..3) of a (h, w) u8 image whose sides are divisible by 2^level, by a
    plain integer tent sum: weights (1, 3, .., 2r-1, 2r-1, .., 3, 1) per axis, r = 2^level, total 4 r^4,
    round half up.  Returns (bytes[h_i - 2, w_i - 2]) for outputs 1 .. n-2 of both axes (complete windows)."""
    r = 1 << level
    k = np.concatenate([np.arange(1, 2 * r, 2), np.arange(2 * r - 1, 0, -2)]).astype(np.int64)   # 2r taps
    h, w = img.shape
    ho, wo = h >> level, w >> level
    a = img.astype(np.int64)
    out = np.zeros((ho - 2, wo - 2), np.int64)
    tot = int(k.sum()) ** 2
    for oy in range(1, ho - 1):
        ys = oy * r - r // 2                                   # first tap: o r + (r - 1)/2 - (r - 1/2) rounded up
        rows = (a[ys:ys + 2 * r, :] * k[:, None]).sum(axis=0)
        for ox in range(1, wo - 1):
            xs = ox * r - r // 2
            out[oy - 1, ox - 1] = (int((rows[xs:xs + 2 * r] * k).sum()) * 2 + tot) // (2 * tot)
    return out.astype(np.uint8)


def pillow_levels(img, levels):
    from PIL import Image
    h, w = img.shape
    im = Image.fromarray(img)
    return [np.asarray(im.resize((w >> i, h >> i), Image.BILINEAR)) if i else img.copy() for i in range(levels)]


def tie_heavy_image(w, h, seed):
    """few grey levels in blocks + sparse impulses: many equal scores, many contrasts exactly on a threshold"""
    rng = np.random.default_rng(seed)
    base = rng.integers(0, 6, ((h + 7) // 8, (w + 7) // 8)) * 40 + 20
    img = np.kron(base, np.ones((8, 8), np.int64))[:h, :w]
    ys, xs = rng.integers(0, h, w * h // 40), rng.integers(0, w, w * h // 40)
    img[ys, xs] += rng.choice([-40, -35, -10, 10, 11, 35, 40], len(xs))
    return np.clip(img, 0, 255).astype(np.uint8)


def pin_images(w, h):
    """>= 5 seeded images: analytic scene, uniform noise, tie-heavy blocks, checkerboard + noise, ramp + impulses"""
    from util import random_image, scene_pair
    rng = np.random.default_rng(77)
    yy, xx = np.mgrid[0:h, 0:w]
    checker = (((xx // 6) + (yy // 6)) % 2) * 90 + 60 + rng.integers(-6, 7, (h, w))
    ramp = (xx * 255 // max(w - 1, 1)).astype(np.int64)
    ramp[rng.integers(0, h, 600), rng.integers(0, w, 600)] += rng.integers(-80, 81, 600)
    return [scene_pair(w, h, 31)[0], random_image(w, h, 32), tie_heavy_image(w, h, 33),
            np.clip(checker, 0, 255).astype(np.uint8), np.clip(ramp, 0, 255).astype(np.uint8)]
