"""CPU: oracle/crate_b.py (tracking core of the reference's second tracker, row f-2 groundwork) against the
reference's OWN unit tests -- the only known-answer vectors the reference ships for any tracker:
feature_tracker/src/feature_tracker/feature_tracking.rs:264-297 (exp_se2 / log_se2),
feature_tracker/src/image_operations.rs:84-94 (pyramid sizes) and :305-368 (bicubic derivatives)."""
import math
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import crate_b as B  # noqa: E402


def relative_eq(a, b, eps, max_rel):
    a, b = np.asarray(a, np.float32), np.asarray(b, np.float32)
    d = np.abs(a - b)
    return bool(np.all((d <= eps) | (d <= max_rel * np.maximum(np.abs(a), np.abs(b)))))


def test_exp_se2_known_answers():
    tol = 1e-6
    T = B.exp_se2([0.0, 0.0, 0.0])
    assert relative_eq(T.matrix()[:2, :2], np.eye(2), tol, tol) and relative_eq(T.t, [0, 0], tol, tol)
    T = B.exp_se2([np.float32(math.pi), 0.0, 0.0])
    assert relative_eq(T.matrix()[:2, :2], -np.eye(2), tol, tol) and relative_eq(T.t, [0, 0], tol, tol)
    T = B.exp_se2([0.0, 1.32, -1.56])
    assert relative_eq(T.matrix()[:2, :2], np.eye(2), tol, tol) and relative_eq(T.t, [1.32, -1.56], tol, tol)


def test_exp_log_round_trips():
    pi = np.float32(math.pi)
    for tw in ([0.0, 0.0, 0.0], [pi / 3, 1.0, 2.5], [1.795, 1.0, 2.5], [1e-8, 1.0, 2.5], [1e-7, 1.0, 2.5], [2e-7, 1.0, 2.5]):
        tw = np.array(tw, np.float32)
        back = B.log_se2(B.exp_se2(tw))
        assert relative_eq(tw, back, 1e-6, 1e-6), (tw, back)


def test_bicubic_derivatives_by_finite_differences():
    delta = np.float32(math.sqrt(np.finfo(np.float32).eps))
    tol = 50.0 * delta
    rng = np.random.default_rng(1)
    for _ in range(50):
        f = rng.standard_normal(4).astype(np.float32)
        t = np.float32(rng.standard_normal())
        fd = (B.bicubic_1d(f, t + delta) - B.bicubic_1d(f, t - delta)) / (2 * delta)
        assert abs(fd - B.d_bicubic_1d(f, t)[1]) <= tol
        assert abs(fd - B.d_bicubic_1d_full(f, t)[2]) <= tol
        for i in range(4):
            f = rng.standard_normal(4).astype(np.float32)
            t = np.float32(rng.standard_normal())
            df = np.zeros(4, np.float32)
            df[i] = delta
            fd = (B.bicubic_1d(f + df, t) - B.bicubic_1d(f - df, t)) / (2 * delta)
            assert abs(B.d_bicubic_1d_full(f, t)[1][i] - fd) <= tol


def test_pyramid_sizes_and_constant_image():
    img = np.full((60, 120), 0.25, np.float32)
    pyr = B.build_image_pyramid(img, 3, 2.0)
    assert [p.shape for p in pyr] == [(60, 120), (30, 60), (15, 30)]
    assert all(np.allclose(p, 0.25, atol=1e-6) for p in pyr)
    assert [p.shape for p in B.build_image_pyramid(np.zeros((45, 77), np.float32), 3, 1.5)] == [(45, 77), (30, 51), (20, 34)]


def _scene(h, w, dx=0.0, dy=0.0):
    y, x = np.mgrid[0:h, 0:w].astype(np.float64)
    x, y = x - dx, y - dy
    v = 0.5 + 0.18 * np.sin(0.21 * x + 0.13 * y) + 0.14 * np.sin(0.09 * x - 0.17 * y + 1.0) + 0.1 * np.cos(0.05 * x * 0 + 0.3 * y)
    return v.astype(np.float32)


def test_sampler_bounds_and_values():
    img = _scene(40, 50)
    assert B.interpolate_bicubic(img, 0.5, 10.0) is None and B.interpolate_bicubic(img, 10.0, 38.0) is None
    assert B.interpolate_bicubic(img, -3.0, 10.0) is None
    assert abs(B.interpolate_bicubic(img, 12.0, 7.0) - img[7, 12]) < 1e-6          # t = 0: the pixel itself
    v, g = B.d_interpolate_bicubic(img, 20.3, 15.6)
    h = 1e-2
    fdx = (B.interpolate_bicubic(img, 20.3 + h, 15.6) - B.interpolate_bicubic(img, 20.3 - h, 15.6)) / (2 * h)
    fdy = (B.interpolate_bicubic(img, 20.3, 15.6 + h) - B.interpolate_bicubic(img, 20.3, 15.6 - h)) / (2 * h)
    assert abs(g[0] - fdx) < 2e-3 and abs(g[1] - fdy) < 2e-3


def test_tracking_recovers_a_known_shift():
    h, w = 96, 128
    a, b = _scene(h, w), _scene(h, w, 1.3, -0.7)
    pa, pb = B.build_image_pyramid(a, 3, 2.0), B.build_image_pyramid(b, 3, 2.0)
    pts = [(40.0, 40.0), (64.0, 50.0), (90.0, 60.0)]
    out = B.track_points(pa, pb, pts, max_iteration=10, lm_lambda=0.1)
    assert sorted(out) == [0, 1, 2]
    for k, T in out.items():
        moved = T.apply(pts[k]) - np.array(pts[k], np.float32)
        assert np.abs(moved - np.array([1.3, -0.7], np.float32)).max() < 0.05, moved


# ---- vectors of the real crate (tools/reference_vectors_b; absent here: no cargo) --------------------------------
VECTORS_B = os.environ.get("RSVIO_REFERENCE_VECTORS_B",
                           os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_vectors_b.txt"))
M32 = np.uint64(0xFFFFFFFF)


def mix(x, y, seed):
    x, y = np.asarray(x, np.uint64), np.asarray(y, np.uint64)
    s = ((x * np.uint64(0x9E3779B1)) & M32) ^ ((y * np.uint64(0x85EBCA77)) & M32) ^ np.uint64((seed * 0xC2B2AE3D) & 0xFFFFFFFF)
    s ^= s >> np.uint64(15)
    s = (s * np.uint64(0x2C1B3C6D)) & M32
    s ^= s >> np.uint64(12)
    s = (s * np.uint64(0x297A2D39)) & M32
    s ^= s >> np.uint64(15)
    return s


def scene_b(w, h, seed):
    y, x = np.mgrid[0:h, 0:w]
    block = (mix(x // 8, y // 8, seed) >> np.uint64(24)) & np.uint64(0xE0)
    noise = mix(x, y, seed ^ 0x5555) >> np.uint64(28)
    return ((block + noise).astype(np.float32) / np.float32(255.0)).astype(np.float32)


def _f32s(tokens):
    return np.array([int(t, 16) for t in tokens], np.uint32).view(np.float32)


def test_scene_b_generator():
    a = scene_b(40, 30, 61)
    assert a.dtype == np.float32 and a.shape == (30, 40) and 0.0 <= a.min() and a.max() <= 1.0
    assert np.array_equal(a, scene_b(40, 30, 61)) and len(np.unique(a)) > 20


def test_crate_b_oracle_matches_the_reference_vectors():
    import pytest
    if not os.path.exists(VECTORS_B):
        pytest.skip("tests/golden/reference_vectors_b.txt not generated (needs cargo: tools/reference_vectors_b)")
    for line in open(VECTORS_B):
        if line.startswith("#") or ":" not in line:
            continue
        head, _, tail = line.partition(":")
        head = head.split()
        if head[0] == "pyr_b":
            w, h, seed, levels = map(int, head[1:5])
            ratio, blur, lev, lw, lh = float(_f32s([head[5]])[0]), int(head[6]), int(head[8]), int(head[9]), int(head[10])
            if blur:
                continue                                   # imageops::blur is not restated yet
            pyr = B.build_image_pyramid(scene_b(w, h, seed), levels, ratio)
            want = _f32s(tail.split()).reshape(lh, lw)
            assert pyr[lev].shape == (lh, lw)
            assert np.abs(pyr[lev] - want).max() <= 1e-6, f"pyramid {w}x{h} level {lev}"
        elif head[0] == "bicubic":
            w, h, seed = map(int, head[1:4])
            x, y, ok = _f32s([head[4]])[0], _f32s([head[5]])[0], int(head[6])
            r = B.d_interpolate_bicubic(scene_b(w, h, seed), x, y)
            assert (r is not None) == bool(ok)
            if ok:
                v, v2, gx, gy = _f32s(tail.split())
                assert abs(r[0] - v) <= 1e-6 and abs(B.interpolate_bicubic(scene_b(w, h, seed), x, y) - v2) <= 1e-6
                assert np.abs(r[1] - np.array([gx, gy], np.float32)).max() <= 1e-5
        elif head[0] == "patch_b":
            w, h, seed = map(int, head[1:4])
            x, y = _f32s([head[4]])[0], _f32s([head[5]])[0]
            d, j, hi = (_f32s(part.split()) for part in tail.split("|"))
            p = B.Patch52(scene_b(w, h, seed), (x, y), 0.1)
            assert np.abs(p.intensities - d).max() <= 1e-6
            assert np.abs(p.J - j.reshape(52, 3)).max() <= 1e-3 * max(1.0, np.abs(j).max())
            assert np.abs(p.Hinv - hi.reshape(3, 3)).max() <= 1e-3 * np.abs(hi).max()


# ---- the rest of crate B in the oracle: blur, fast_blur, Shi-Tomasi, NMS, local_maxima, FeatureTracker ---------------
def _texture(h, w, seed, dx=0.0, dy=0.0):
    rng = np.random.default_rng(seed)
    y, x = np.mgrid[0:h, 0:w].astype(np.float64)
    x, y = x - dx, y - dy
    v = 0.5 + 0.12 * np.sin(0.21 * x + 0.13 * y) + 0.1 * np.sin(0.09 * x - 0.17 * y + 1.0)
    for _ in range(max(6, h * w // 900)):
        cx, cy, s, a = rng.uniform(8, w - 8), rng.uniform(8, h - 8), rng.uniform(1.5, 3.5), rng.uniform(-0.3, 0.3)
        v += a * np.exp(-((x - cx) ** 2 + (y - cy) ** 2) / (2 * s * s))
    return np.clip(v, 0.0, 1.0).astype(np.float32)


def test_blur_properties():
    flat = np.full((20, 30), 0.4, np.float32)
    assert np.abs(B.blur(flat, 2.0) - 0.4).max() < 1e-6                  # weights are normalised, also at the border
    img = np.zeros((41, 41), np.float32)
    img[20, 20] = 1.0
    b = B.blur(img, 2.0)
    assert abs(b.sum() - 1.0) < 1e-5 and b[20, 20] == b.max()
    assert np.count_nonzero(b[20]) == 9 and np.count_nonzero(b[:, 20]) == 9   # support 2 sigma: taps -4 .. 4
    assert np.allclose(b, b.T) and np.allclose(b, b[::-1, ::-1])
    # against OpenCV's Gaussian (different truncation: 2 sigma here), interior only
    import cv2
    tex = _texture(60, 80, 1)
    ref = cv2.GaussianBlur(tex, (9, 9), 2.0, borderType=cv2.BORDER_REPLICATE)
    assert np.abs(B.blur(tex, 2.0) - ref)[6:-6, 6:-6].max() < 2e-3          # cv2 with the same 9 taps


def test_fast_blur_boxes_and_properties():
    assert B.boxes_for_gauss(6.0) == [11, 11, 13] and B.boxes_for_gauss(2.0) == [3, 3, 5] and B.boxes_for_gauss(1.0) == [1, 1, 3]
    flat = np.full((17, 23), 0.3, np.float32)
    assert np.abs(B.fast_blur(flat, 6.0) - 0.3).max() < 1e-6
    img = np.zeros((31, 31), np.float32)
    img[15, 15] = 1.0
    b = B.fast_blur(img, 2.0)
    assert b.shape == img.shape and abs(b.sum() - 1.0) < 1e-5 and b[15, 15] == b.max()
    neg = np.full((9, 9), -0.5, np.float32)
    assert (B.fast_blur(neg, 1.0) == 0).all()                             # values are clamped to the sample range [0, 1]


def test_score_map_and_nms():
    flat = np.full((30, 30), 0.5, np.float32)
    assert (B.shi_tomasi_score(flat, 2.0) == 0).all()
    img = B.blur(_texture(60, 80, 2), 1.5)
    s = B.shi_tomasi_score(img, 2.0)
    assert s.shape == img.shape and s.min() >= 0 and s.max() > 0.05
    for thr in (0.0, 0.05):
        a, b = B.suppress_non_maximum(s, 1, np.float32(thr)), B.suppress_non_maximum_r1(s, np.float32(thr))
        assert a == b and len(a) > 5
        for x, y, v in a[:20]:                                             # every kept pixel is the 3x3 maximum
            assert v == s[max(y - 1, 0):y + 2, max(x - 1, 0):x + 2].max() and v >= thr
    q = (np.round(s * 40) / 40).astype(np.float32)                         # many exact ties
    assert B.suppress_non_maximum(q, 1, np.float32(0.0)) == B.suppress_non_maximum_r1(q, np.float32(0.0))


def test_local_maxima_follows_imageproc():
    items = [(10, 10, 1.0, "a"), (12, 11, 2.0, "b"), (30, 10, 1.0, "c"), (10, 13, 1.0, "d"), (50, 40, 0.5, "e")]
    assert [t[3] for t in B.local_maxima(items, 3)] == ["c", "b", "e"]    # a and d lose to b; order is (y, x)
    tie = [(5, 5, 1.0, "p"), (6, 5, 1.0, "q"), (20, 9, 1.0, "z")]
    assert [t[3] for t in B.local_maxima(tie, 2)] == ["p", "z"]            # ties go to the lexicographically first
    # the last occupied row is never searched: a stronger corner there does not suppress its neighbours
    last = [(5, 5, 1.0, "w"), (6, 6, 9.0, "s")]
    assert [t[3] for t in B.local_maxima(last, 3)] == ["w", "s"]
    assert B.local_maxima([], 5) == []


def test_add_points_and_tracked_features():
    img = B.blur(_texture(96, 128, 4), 1.5)
    first = B.add_points(img, [], 0.05, 7, 2.0)
    assert len(first) >= 5 and first == sorted(first, key=lambda c: (c[1], c[0]))
    assert all(7 <= x < 128 - 7 and 7 <= y < 96 - 7 for x, y in first)
    for i, (x0, y0) in enumerate(first):                                   # kept corners are more than min_dist apart
        for x1, y1 in first[i + 1:]:
            assert max(abs(x0 - x1), abs(y0 - y1)) > 7 or y1 == first[-1][1]
    again = B.add_points(img, [(x + 0.2, y - 0.3) for x, y in first], 0.05, 7, 2.0)
    assert not set(again) & set(first)                                     # occupied neighbourhoods yield nothing new


def test_feature_tracker_sequence():
    h, w = 96, 128
    ft = B.FeatureTracker(nlevels=3, preprocessing_blur_sigma=1.5, detection_threshold=0.02, detection_min_dist=9,
                          detection_blur=2.0, optical_flow_max_iter=12)
    assert ft.get_pyramid() is None
    f0 = ft.process_frame(_texture(h, w, 21))
    assert len(f0) >= 6 and [f[0] for f in f0] == list(range(len(f0)))
    f1 = ft.process_frame(_texture(h, w, 21, dx=1.2, dy=-0.7))
    p0 = {i: (x, y) for i, x, y in f0}
    kept = [(i, x, y) for i, x, y in f1 if i in p0]
    assert len(kept) >= len(f0) // 2
    flow = np.array([[x - p0[i][0], y - p0[i][1]] for i, x, y in kept])
    assert np.abs(np.median(flow, axis=0) - np.array([1.2, -0.7])).max() < 0.1
    new_ids = [i for i, _, _ in f1 if i not in p0]
    assert new_ids == list(range(len(f0), len(f0) + len(new_ids)))        # fresh ids, after the tracked features
    assert [p.shape for p in ft.get_pyramid()] == [(96, 128), (48, 64), (24, 32)]


def test_lssd_patch():
    img = B.blur(_texture(64, 64, 7), 1.0)
    p = B.Patch52Lssd(img, (30.3, 31.6), 0.1)
    r = p.residuals(img, B.Iso2())
    mean = np.float32(p.intensities.sum() / 52)
    assert np.allclose(r, p.intensities / mean - p.intensities, atol=1e-5)   # the template stays raw, as in the reference
    T = B.track_point([img], [(img * np.float32(0.7)).astype(np.float32)], (30.3, 31.6), 15, 0.1, "lssd")
    assert T is not None


def test_crate_b_image_ops_against_independent_implementations():
    """Third-party twins of the restated crate arithmetic (scipy / OpenCV / numpy.linalg share no code with the oracle):
    they pin the ALGORITHMS as restated (box sizes, edge replication, kernel taps, the eigenvalue formula), not the
    crates' own rounding order."""
    import cv2
    from scipy import ndimage
    rng = np.random.default_rng(3)
    img = B.blur(_texture(72, 96, 5), 1.0)
    # fast_blur == three passes of a replicated-edge box filter of sizes boxes_for_gauss(sigma), both axes
    for sigma in (1.0, 2.0, 6.0):
        plane = rng.uniform(0.0, 1.0, (50, 64)).astype(np.float32)
        want = plane.astype(np.float64)
        for size in B.boxes_for_gauss(sigma):
            want = ndimage.uniform_filter1d(want, size, axis=1, mode="nearest")
            want = ndimage.uniform_filter1d(want, size, axis=0, mode="nearest")
        assert np.abs(B.fast_blur(plane, sigma) - want).max() < 2e-6, sigma
    # the score is 1000 x the smaller eigenvalue of the blurred structure tensor
    p = np.pad(img, 1, mode="edge")
    dx, dy = p[1:-1, 2:] - p[1:-1, :-2], p[2:, 1:-1] - p[:-2, 1:-1]
    a, b, c = (B.fast_blur((u * v).astype(np.float32), 2.0).astype(np.float64) for u, v in ((dx, dx), (dy, dy), (dx, dy)))
    lam = np.linalg.eigvalsh(np.stack([np.stack([a, c], -1), np.stack([c, b], -1)], -2))[..., 0]
    assert np.abs(B.shi_tomasi_score(img, 2.0) - 1000.0 * np.abs(lam)).max() < 2e-3 * max(1.0, 1000.0 * np.abs(lam).max())
    # blur == the normalised Gaussian taps -4 .. 4 (sigma 2) applied separably; interior only (truncated windows are
    # renormalised at the border, OpenCV replicates instead)
    k = np.exp(-np.arange(-4, 5, dtype=np.float64) ** 2 / (2 * 2.0 ** 2))
    k /= k.sum()
    tex = _texture(60, 80, 1)
    ref = cv2.sepFilter2D(tex.astype(np.float64), -1, k, k, borderType=cv2.BORDER_REPLICATE)
    assert np.abs(B.blur(tex, 2.0) - ref)[4:-4, 4:-4].max() < 1e-6
