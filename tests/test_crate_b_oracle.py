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
