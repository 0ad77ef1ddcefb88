"""GPU: first CUDA slice of row f-2 (crate B's tracking core, csrc/trackb.cu) against oracle/crate_b.py on the same
f32 pyramids.  Tolerance: isometries to 1e-3 px / 1e-4 rad-ish (the 52-term sums are shuffle trees on the device and
BLAS dot products in the numpy oracle), identical status except on a threshold edge."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import crate_b as B  # noqa: E402

pytestmark = pytest.mark.gpu


def scene(h, w, dx=0.0, dy=0.0, rot=0.0):
    y, x = np.mgrid[0:h, 0:w].astype(np.float64)
    cx, cy = w / 2.0, h / 2.0
    xr = np.cos(rot) * (x - cx) + np.sin(rot) * (y - cy) + cx - dx
    yr = -np.sin(rot) * (x - cx) + np.cos(rot) * (y - cy) + cy - dy
    v = 0.5 + 0.18 * np.sin(0.21 * xr + 0.13 * yr) + 0.14 * np.sin(0.09 * xr - 0.17 * yr + 1.0) + 0.1 * np.cos(0.3 * yr)
    return v.astype(np.float32)


def oracle_track(pa, pb, pts, max_it, lam):
    """(T4, status) in the layout of rsvio_b_track_points"""
    T4 = np.zeros((len(pts), 4), np.float32)
    st = np.zeros(len(pts), np.int32)
    for k, p in enumerate(pts):
        T0 = B.track_point(pa, pb, p, max_it, lam)
        if T0 is None:
            st[k] = 1
            T4[k] = (1, 0, 0, 0)
            continue
        T4[k] = (T0.c, T0.s, T0.t[0], T0.t[1])
        p1 = T0.apply(p)
        T1 = B.track_point(pb, pa, p1, max_it, lam)
        if T1 is None:
            st[k] = 2
            continue
        back = T1.apply(p1)
        if not np.float32(np.linalg.norm(np.array(p, np.float32) - back)) < np.float32(2.0):
            st[k] = 3
    return T4, st


def test_crate_b_tracking_core_matches_oracle(rsvio):
    h, w = 96, 128
    a, b = scene(h, w), scene(h, w, 1.3, -0.7, 0.01)
    pa, pb = B.build_image_pyramid(a, 3, 2.0), B.build_image_pyramid(b, 3, 2.0)
    rng = np.random.default_rng(3)
    pts = np.concatenate([rng.uniform([20, 20], [w - 20, h - 20], (24, 2)),       # interior
                          [[3.0, 3.0], [w - 4.0, h - 5.0], [8.0, 50.0], [64.0, 6.0], [0.0, 0.0]]]).astype(np.float32)
    want_T, want_st = oracle_track(pa, pb, pts, 10, 0.1)
    got_T, got_st = rsvio.b_track_points(pa, pb, pts, 10, 0.1)
    assert (want_st[:24] == 0).all()                                               # the scene is trackable
    same = got_st == want_st
    assert same.mean() >= 0.9, (got_st, want_st)
    ok = same & (want_st != 1)
    moved_g = np.stack([got_T[:, 0] * pts[:, 0] - got_T[:, 1] * pts[:, 1] + got_T[:, 2],
                        got_T[:, 1] * pts[:, 0] + got_T[:, 0] * pts[:, 1] + got_T[:, 3]], 1)
    moved_w = np.stack([want_T[:, 0] * pts[:, 0] - want_T[:, 1] * pts[:, 1] + want_T[:, 2],
                        want_T[:, 1] * pts[:, 0] + want_T[:, 0] * pts[:, 1] + want_T[:, 3]], 1)
    good = ok & (want_st == 0)
    assert np.abs(moved_g[good] - moved_w[good]).max() < 1e-2
    assert np.median(np.abs(moved_g[good] - moved_w[good])) < 1e-3
    # and both recover the synthetic motion of the interior points
    flow = moved_g[:24] - pts[:24]
    assert np.abs(np.median(flow, axis=0) - np.array([1.3, -0.7])).max() < 0.1


def test_crate_b_pyramid_matches_oracle(rsvio):
    """f32 chained Triangle resize: same operations in the same order as the numpy restatement -> same bits"""
    rng = np.random.default_rng(8)
    for (h, w, levels, ratio) in ((60, 120, 3, 2.0), (75, 97, 4, 2.0), (64, 96, 3, 1.5), (33, 41, 2, 3.0)):
        img = rng.uniform(0.0, 1.0, (h, w)).astype(np.float32)
        want = B.build_image_pyramid(img, levels, ratio)
        got = rsvio.b_build_pyramid(img, levels, ratio)
        assert [g.shape for g in got] == [x.shape for x in want]
        for l in range(levels):
            assert np.abs(got[l] - want[l]).max() <= 1e-7, (h, w, ratio, l, np.abs(got[l] - want[l]).max())
    # the device pyramid feeds the device tracker
    a, b = scene(96, 128), scene(96, 128, 1.3, -0.7)
    pa, pb = rsvio.b_build_pyramid(a, 3, 2.0), rsvio.b_build_pyramid(b, 3, 2.0)
    pts = np.array([[40.0, 40.0], [64.0, 50.0], [90.0, 60.0]], np.float32)
    T, st = rsvio.b_track_points(pa, pb, pts, 10, 0.1)
    flow = np.stack([T[:, 0] * pts[:, 0] - T[:, 1] * pts[:, 1] + T[:, 2],
                     T[:, 1] * pts[:, 0] + T[:, 0] * pts[:, 1] + T[:, 3]], 1) - pts     # T p - p, not T's translation
    assert (st == 0).all() and np.abs(flow - np.array([1.3, -0.7])).max() < 0.05
