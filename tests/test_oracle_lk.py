"""CPU: the oracle's LK pieces -- reference known-answer vectors for exp_se2, the numpy
restatement of Pattern52::new, analytic ground truth on synthetic sub-pixel shifts, and
failure-path properties.  PARITY UNPINNED for the tracker itself (no reference fixtures)."""
import math

import numpy as np
import pytest

from oracle import np_crosscheck as npx
from util import grid_points, identity_T6, scene_pair


def test_se2_exp_known_answers_from_reference(oracle):
    # reference feature_tracker/src/feature_tracker/feature_tracking.rs:264-291 (twist order there is
    # [theta, vx, vy]; tracker A uses [vx, vy, theta], image_utilities.rs:83,103-104)
    m = oracle.se2_exp([0.0, 0.0, 0.0])
    assert np.allclose(m, np.eye(3), atol=1e-6)
    m = oracle.se2_exp([0.0, 0.0, math.pi])
    assert np.allclose(m, [[-1, 0, 0], [0, -1, 0], [0, 0, 1]], atol=1e-6)
    m = oracle.se2_exp([1.32, -1.56, 0.0])
    assert np.allclose(m, [[1, 0, 1.32], [0, 1, -1.56], [0, 0, 1]], atol=1e-6)


def test_se2_exp_closed_form(oracle):
    rng = np.random.default_rng(0)
    for _ in range(200):
        vx, vy, th = rng.uniform(-3, 3), rng.uniform(-3, 3), rng.uniform(-3, 3)
        if abs(th) < 0.1:  # the reference's f32 (1 - cos t) / t cancels badly for small t (kept as is)
            continue
        m = oracle.se2_exp([vx, vy, th]).astype(np.float64)
        s, c = math.sin(th), math.cos(th)
        V = np.array([[s / th, -(1 - c) / th], [(1 - c) / th, s / th]])
        assert np.allclose(m[:2, :2], [[c, -s], [s, c]], atol=1e-6)
        assert np.allclose(m[:2, 2], V @ [vx, vy], atol=1e-5)
    # small-angle branch
    m = oracle.se2_exp([0.5, -0.25, 1e-9])
    assert np.allclose(m[:2, 2], [0.5, -0.25], atol=1e-6)


def test_pattern52_matches_numpy_restatement(oracle):
    img = scene_pair(160, 120, 21)[0]
    for (px, py) in [(80.3, 60.7), (20.0, 20.0), (3.2, 50.0), (155.9, 117.2), (100.5, 4.4)]:
        p = oracle.pattern52_new(img, px, py)
        ref = npx.pattern52_new(img, px, py)
        data = np.array(p.data[:], np.float32)
        assert np.array_equal(data, ref["data"]) or (np.isnan(data).any() and ref["n"] == 0)
        if ref["n"] > 0:
            assert np.float32(p.mean) == ref["mean"]


def test_all_black_and_flat_patches_are_invalid(oracle):
    assert oracle.pattern52_new(np.zeros((64, 64), np.uint8), 32.0, 32.0).valid == 0
    assert oracle.pattern52_new(np.full((64, 64), 200, np.uint8), 32.0, 32.0).valid == 0  # singular Hessian


def test_tracks_known_subpixel_shift(oracle):
    w, h, levels = 320, 240, 3
    shift = (1.37, -0.62)
    a, b = scene_pair(w, h, 31, shift)
    pa, pb = oracle.build_pyramid(a, levels), oracle.build_pyramid(b, levels)
    xs, ys, _ = oracle.detect_key_points(a, 20, np.zeros((0, 2), np.uint32))
    T0 = identity_T6(np.stack([xs, ys], axis=1))
    T1, status, iters, margins = oracle.track_points(pa, pb, w, h, levels, T0, 20, 0.01, nthreads=4)
    ok = status == 0
    assert ok.mean() > 0.8
    d = T1[ok][:, [2, 5]] - T0[ok][:, [2, 5]]
    err = np.abs(d - np.array(shift, np.float32))
    assert np.median(err) < 0.05 and np.percentile(err, 95) < 0.3
    assert (iters[ok] > 0).all()
    # results do not depend on the thread count
    T1b, st_b, _, _ = oracle.track_points(pa, pb, w, h, levels, T0, 20, 0.01, nthreads=1)
    assert np.array_equal(T1, T1b) and np.array_equal(status, st_b)


def test_forward_backward_rejects_changed_content(oracle):
    w, h, levels = 200, 160, 3
    a, _ = scene_pair(w, h, 41)
    b = scene_pair(w, h, 4242)[0]  # unrelated texture
    pa, pb = oracle.build_pyramid(a, levels), oracle.build_pyramid(b, levels)
    T0 = identity_T6(grid_points(w, h, 200, 5))
    _, status, _, _ = oracle.track_points(pa, pb, w, h, levels, T0, 20, 0.01)
    assert (status != 0).mean() > 0.9


def test_points_near_border_fail_cleanly(oracle):
    w, h, levels = 128, 96, 4
    a, b = scene_pair(w, h, 51)
    pa, pb = oracle.build_pyramid(a, levels), oracle.build_pyramid(b, levels)
    xy = np.array([[0.0, 0.0], [1.0, 50.0], [127.0, 95.0], [64.0, 1.5], [-5.0, 10.0], [1e9, 3.0]], np.float32)
    T1, status, _, _ = oracle.track_points(pa, pb, w, h, levels, identity_T6(xy), 10, 0.01)
    assert (status != 0).all()
    assert np.isfinite(T1[:4]).all()


def test_stereo_tracker_sequence(oracle):
    from rsvio_b200.synth import StereoSequence
    seq = StereoSequence("C1", w=376, h=240, grid=40)
    tr = oracle.StereoTracker(3, 40, 20, 0.01, nthreads=4)
    first = None
    for t in range(4):
        l, r = seq.frame(t)
        tr.process_frame(l, r)
        ids0, T0 = tr.get_tracks(0)
        ids1, T1 = tr.get_tracks(1)
        assert np.array_equal(np.sort(ids0), ids0) and len(set(ids0.tolist())) == len(ids0)
        if first is None:
            first = dict(zip(ids0.tolist(), T0[:, [2, 5]].copy()))
            assert np.array_equal(ids0, ids1)          # ids are assigned in stereo pairs
            assert np.array_equal(ids0, np.arange(len(ids0), dtype=np.uint64))
        m1 = dict(zip(ids1.tolist(), T1))
        disp = np.array([m1[i][[2, 5]] - T[[2, 5]] for i, T in zip(ids0.tolist(), T0) if i in m1])
        assert np.abs(np.median(disp, axis=0) - np.array(seq.params["disparity"])).max() < 0.1
        err = [T[[2, 5]] - seq.true_left(first[i], t) for i, T in zip(ids0.tolist(), T0) if i in first]
        assert np.median(np.abs(err)) < 0.1
    tracks, frames = tr.counters()
    assert frames == 4 and tracks > 0
    ids0, _ = tr.get_tracks(0)
    tr.remove_ids(ids0[:5])
    assert len(tr.get_tracks(0)[0]) == len(ids0) - 5
