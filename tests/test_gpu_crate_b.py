"""GPU: row f-2 (crate B: csrc/trackb.cu, detectb.cu, apib.cu) against oracle/crate_b.py.  Bars: tracked positions
within 1e-3 px and identical status (the 52-term sums are shuffle trees on the device and BLAS dot products in the
numpy oracle); blur, chained pyramid, score map and corner lists bit-identical (same f32 operations in the same
order)."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import crate_b as B  # noqa: E402

pytestmark = pytest.mark.gpu


def record(stats):
    """append the parity statistics of a comparison to gpurun_out/parity_stats.jsonl (if the directory exists)"""
    import json
    d = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
    if os.path.isdir(d):
        with open(os.path.join(d, "parity_stats.jsonl"), "a") as f:
            f.write(json.dumps(stats) + "\n")


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
    moved_g = np.stack([got_T[:, 0] * pts[:, 0] - got_T[:, 1] * pts[:, 1] + got_T[:, 2],
                        got_T[:, 1] * pts[:, 0] + got_T[:, 0] * pts[:, 1] + got_T[:, 3]], 1)
    moved_w = np.stack([want_T[:, 0] * pts[:, 0] - want_T[:, 1] * pts[:, 1] + want_T[:, 2],
                        want_T[:, 1] * pts[:, 0] + want_T[:, 0] * pts[:, 1] + want_T[:, 3]], 1)
    good = (got_st == want_st) & (want_st == 0)
    err = np.abs(moved_g[good] - moved_w[good])
    record({"tag": "crate B track_points 128x96 L3", "n": int(len(pts)), "status_equal": int((got_st == want_st).sum()),
            "tracked": int(good.sum()), "max_err_px": float(err.max()), "median_err_px": float(np.median(err))})
    # same bar as tracker A: identical status, positions within 1e-3 px (the interior points; the five border points
    # exercise the out-of-bound exits and must agree in status)
    assert (got_st == want_st).all(), (got_st, want_st)
    assert err.max() < 1e-3, err.max()
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


# ---- row f-2, the rest of crate B: blur, LSSD, Shi-Tomasi detection, stateful FeatureTracker ----------------------------
def textured(h, w, seed, dx=0.0, dy=0.0):
    """smooth texture + a few bright blobs (corner-like), values in [0, 1], shifted by (dx, dy)"""
    rng = np.random.default_rng(seed)
    y, x = np.mgrid[0:h, 0:w].astype(np.float64)
    x, y = x - dx, y - dy
    v = 0.5 + 0.12 * np.sin(0.21 * x + 0.13 * y) + 0.1 * np.sin(0.09 * x - 0.17 * y + 1.0)
    for _ in range(max(6, h * w // 900)):
        cx, cy, s, a = rng.uniform(8, w - 8), rng.uniform(8, h - 8), rng.uniform(1.5, 3.5), rng.uniform(-0.3, 0.3)
        v += a * np.exp(-((x - cx) ** 2 + (y - cy) ** 2) / (2 * s * s))
    return np.clip(v, 0.0, 1.0).astype(np.float32)


def test_crate_b_blur_is_bit_identical_to_the_oracle(rsvio):
    rng = np.random.default_rng(5)
    for (h, w, sigma) in ((61, 83, 2.0), (40, 50, 0.8), (33, 37, 3.3)):
        img = rng.uniform(0.0, 1.0, (h, w)).astype(np.float32)
        want, got = B.blur(img, sigma), rsvio.b_blur(img, sigma)
        assert np.array_equal(got, want), (h, w, sigma, np.abs(got - want).max())
    flat = np.full((24, 31), 0.25, np.float32)
    assert np.abs(rsvio.b_blur(flat, 2.0) - 0.25).max() < 1e-6


def test_crate_b_score_map_is_bit_identical_to_the_oracle(rsvio):
    for (h, w, seed, dblur) in ((72, 96, 1, 6.0), (65, 51, 2, 2.0), (48, 160, 3, 1.0)):
        img = B.blur(textured(h, w, seed), 2.0)
        want, got = B.shi_tomasi_score(img, dblur), rsvio.b_shi_tomasi_score(img, dblur)
        assert np.array_equal(got, want), (h, w, dblur, np.abs(got - want).max())
    assert np.abs(rsvio.b_shi_tomasi_score(np.full((40, 40), 0.5, np.float32), 2.0)).max() == 0.0


def test_crate_b_add_points_matches_the_oracle(rsvio):
    for (h, w, seed, thr, md, dblur) in ((96, 128, 4, 0.05, 7, 2.0), (120, 90, 5, 0.02, 15, 3.0), (64, 64, 6, 0.0, 3, 1.0)):
        img = B.blur(textured(h, w, seed), 1.5)
        first = B.add_points(img, [], thr, md, dblur)
        got = rsvio.b_add_points(img, np.zeros((0, 2), np.float32), thr, md, dblur)
        assert len(first) >= 3, (h, w, len(first))
        assert [tuple(int(v) for v in r) for r in got] == [tuple(c) for c in first]
        # tracked features (some of the first corners, moved by sub-pixel amounts, plus one on the last row) suppress
        # their neighbourhood
        tracked = [(x + 0.3, y - 0.4) for x, y in first[::2]] + [(w - 6.2, h - 1.0)]
        want = B.add_points(img, tracked, thr, md, dblur)
        got = rsvio.b_add_points(img, np.array(tracked, np.float32), thr, md, dblur)
        assert [tuple(int(v) for v in r) for r in got] == [tuple(c) for c in want]
        assert len(want) < len(first)


def test_crate_b_lssd_cost_matches_the_oracle(rsvio):
    h, w = 96, 128
    a, b = scene(h, w), (scene(h, w, 1.1, -0.6, 0.008) * np.float32(0.8)).astype(np.float32)   # gain change
    pa, pb = B.build_image_pyramid(a, 3, 2.0), B.build_image_pyramid(b, 3, 2.0)
    rng = np.random.default_rng(11)
    pts = rng.uniform([24, 24], [w - 24, h - 24], (16, 2)).astype(np.float32)
    got_T, got_st = rsvio.b_track_points_cost(pa, pb, pts, 10, 0.1, rsvio.B_COST_LSSD)
    n_ok = 0
    for k, p in enumerate(pts):
        T0 = B.track_point(pa, pb, p, 10, 0.1, "lssd")
        if T0 is None:
            assert got_st[k] == 1
            continue
        q = T0.apply(p)
        g = np.array([got_T[k, 0] * p[0] - got_T[k, 1] * p[1] + got_T[k, 2], got_T[k, 1] * p[0] + got_T[k, 0] * p[1] + got_T[k, 3]])
        if got_st[k] != 1:
            assert np.abs(g - q).max() < 1e-2, (k, g, q)
            n_ok += 1
    assert n_ok >= 12
    # SSD through the same entry point equals the original one
    T_a, st_a = rsvio.b_track_points_cost(pa, pa, pts, 10, 0.1, rsvio.B_COST_SSD)
    T_b, st_b = rsvio.b_track_points(pa, pa, pts, 10, 0.1)
    assert np.array_equal(T_a, T_b) and np.array_equal(st_a, st_b)


def test_crate_b_stateful_tracker_matches_the_oracle(rsvio):
    """FeatureTracker::process_frame over a short sequence: same ids in the same order, positions within 1e-3 px
    (tracked) / identical (new corners)"""
    h, w = 120, 160
    kw = dict(nlevels=3, ratio=2.0, preprocessing_blur=True, preprocessing_blur_sigma=1.5, detection_threshold=0.02,
              detection_min_dist=9, detection_blur=2.0, optical_flow_max_iter=12, optical_flow_lm_lambda=0.1)
    ref = B.FeatureTracker(**kw)
    gpu = rsvio.FeatureTrackerB(w, h, **kw)
    assert gpu.get_pyramid() is None
    stats = []
    for t in range(4):
        img = textured(h, w, 21, dx=1.2 * t, dy=-0.7 * t)
        want = ref.process_frame(img)
        got = gpu.process_frame(img, t)
        ids_w, ids_g = [f[0] for f in want], [f[0] for f in got]
        assert ids_g == ids_w, (t, len(ids_g), len(ids_w), sorted(set(ids_w) ^ set(ids_g))[:10])
        d = np.abs(np.array([[f[1], f[2]] for f in got]) - np.array([[f[1], f[2]] for f in want]))
        assert d.max() < 1e-3, (t, d.max())
        stats.append((t, len(ids_g), gpu.n_new, float(d.max())))
        for lg, lw in zip(gpu.get_pyramid(), ref.get_pyramid()):
            assert np.array_equal(lg, lw)
    record({"tag": "crate B FeatureTracker 160x120 L3, 4 frames", "per_frame": [list(x) for x in stats]})
    assert stats[0][1] >= 10 and stats[0][1] == stats[0][2]            # first frame: everything is new
    assert stats[-1][1] - stats[-1][2] >= stats[0][1] // 2             # most features survive the sequence
    gpu.close()


def test_crate_b_full_size_frame(rsvio):
    """752x480 with the reference's config.yaml settings: detection runs, tracked features follow a known shift"""
    h, w = 480, 752
    gpu = rsvio.FeatureTrackerB(w, h, detection_threshold=0.02)
    f0 = gpu.process_frame(textured(h, w, 33))
    f1 = gpu.process_frame(textured(h, w, 33, dx=2.0, dy=1.0))
    assert len(f0) >= 50
    p0 = {i: (x, y) for i, x, y in f0}
    flow = np.array([[x - p0[i][0], y - p0[i][1]] for i, x, y in f1 if i in p0])
    assert len(flow) >= len(f0) // 2
    assert np.abs(np.median(flow, axis=0) - np.array([2.0, 1.0])).max() < 0.1
    assert gpu.last_frame_ms() > 0


def test_crate_b_edge_cases(rsvio):
    """flat image -> no features and nothing to track; one level / no blur / LSSD configurations against the oracle;
    a full feature table drops new corners but still returns the frame"""
    h, w = 96, 128
    flat = rsvio.FeatureTrackerB(w, h, nlevels=3, detection_min_dist=7, detection_blur=2.0)
    assert flat.process_frame(np.full((h, w), 0.5, np.float32)) == [] and flat.n_new == 0
    assert flat.process_frame(np.full((h, w), 0.25, np.float32)) == []
    flat.close()
    for kw in (dict(nlevels=1, preprocessing_blur=False), dict(nlevels=2, preprocessing_blur=False, ratio=1.5),
               dict(nlevels=3, preprocessing_blur_sigma=0.8)):
        cfg = dict(ratio=2.0, preprocessing_blur=True, preprocessing_blur_sigma=1.5, detection_threshold=0.02,
                   detection_min_dist=9, detection_blur=2.0, optical_flow_max_iter=8, optical_flow_lm_lambda=0.1)
        cfg.update(kw)
        ref, gpu = B.FeatureTracker(**cfg), rsvio.FeatureTrackerB(w, h, **cfg)
        for t in range(2):
            img = textured(h, w, 9, dx=0.8 * t, dy=0.5 * t)
            want, got = ref.process_frame(img), gpu.process_frame(img, t)
            assert [f[0] for f in got] == [f[0] for f in want], (kw, t)
            d = np.abs(np.array([[f[1], f[2]] for f in got]) - np.array([[f[1], f[2]] for f in want]))
            assert len(got) >= 5 and d.max() < 1e-3, (kw, t, d.max())
        gpu.close()
    # LSSD through the stateful handle (the reference's LSSD compares mean-normalised target samples with a RAW
    # template, so it loses most features -- what matters is that the device loses the same ones as the oracle)
    cfg = dict(nlevels=3, preprocessing_blur_sigma=1.5, detection_threshold=0.02, detection_min_dist=9, detection_blur=2.0,
               optical_flow_max_iter=8)
    ref = B.FeatureTracker(matching_cost="lssd", **cfg)
    lssd = rsvio.FeatureTrackerB(w, h, matching_cost=rsvio.B_COST_LSSD, **cfg)
    for t in range(2):
        img = (textured(h, w, 9, dx=0.6 * t) * np.float32(1.0 - 0.15 * t)).astype(np.float32)
        want, got = ref.process_frame(img), lssd.process_frame(img, t)
        assert [f[0] for f in got] == [f[0] for f in want], t
        d = np.abs(np.array([[f[1], f[2]] for f in got]) - np.array([[f[1], f[2]] for f in want]))
        assert d.max() < 1e-3, (t, d.max())
    lssd.close()
    # capacity: a table of 4 features
    small = rsvio.FeatureTrackerB(w, h, nlevels=2, detection_threshold=0.02, detection_min_dist=9, detection_blur=2.0, max_features=4)
    with pytest.raises(rsvio.RsvioError) as ei:
        small.process_frame(textured(h, w, 9))
    assert ei.value.status == -3                      # RSVIO_ERR_CAPACITY: the frame was processed, excess corners dropped
    assert len(small.features()) == 4
    small.close()


def test_crate_b_track_points_on_random_points(rsvio):
    """96 random points (interior, near the border, outside) on a rotating, shifting scene: identical status, < 1e-3 px"""
    h, w = 120, 160
    a, b = scene(h, w), scene(h, w, -1.6, 0.9, -0.012)
    pa, pb = B.build_image_pyramid(a, 3, 2.0, True, 1.0), B.build_image_pyramid(b, 3, 2.0, True, 1.0)
    rng = np.random.default_rng(21)
    pts = np.concatenate([rng.uniform([-4, -4], [w + 4, h + 4], (72, 2)), rng.uniform([0, 0], [w, 12], (12, 2)),
                          rng.uniform([w - 12, 0], [w, h], (12, 2))]).astype(np.float32)
    want_T, want_st = oracle_track(pa, pb, pts, 12, 0.1)
    got_T, got_st = rsvio.b_track_points(pa, pb, pts, 12, 0.1)

    def moved(T):
        return np.stack([T[:, 0] * pts[:, 0] - T[:, 1] * pts[:, 1] + T[:, 2], T[:, 1] * pts[:, 0] + T[:, 0] * pts[:, 1] + T[:, 3]], 1)
    good = (got_st == want_st) & (want_st == 0)
    err = np.abs(moved(got_T)[good] - moved(want_T)[good])
    record({"tag": "crate B track_points random 160x120 L3", "n": int(len(pts)), "status_equal": int((got_st == want_st).sum()),
            "tracked": int(good.sum()), "status_hist": np.bincount(want_st, minlength=4).tolist(),
            "max_err_px": float(err.max()) if good.any() else 0.0})
    assert (got_st == want_st).all(), (np.flatnonzero(got_st != want_st), got_st[got_st != want_st], want_st[got_st != want_st])
    assert good.sum() >= 40 and err.max() < 1e-3
    assert (want_st != 0).sum() >= 5          # the failure exits are exercised too
