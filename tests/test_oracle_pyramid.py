"""CPU: the C++ oracle's image pyramid against the independent numpy restatement and against
properties of image 0.25's Triangle resize.  PARITY UNPINNED:
the reference holds no golden vectors for tracker A."""
import numpy as np
import pytest

from oracle import np_crosscheck as npx
from util import random_image, scene_pair


@pytest.mark.parametrize("w,h,levels", [(752, 480, 6), (96, 80, 4), (101, 67, 5), (64, 64, 6), (40, 33, 3)])
def test_pyramid_matches_numpy_restatement(oracle, w, h, levels):
    for seed, img in ((1, random_image(w, h, 1)), (2, scene_pair(w, h, 2)[0])):
        flat = oracle.build_pyramid(img, levels)
        got = oracle.split_pyramid(flat, w, h, levels)
        ref = npx.build_pyramid(img, levels)
        for i, (a, b) in enumerate(zip(got, ref)):
            assert a.shape == b.shape == (h >> i, w >> i)
            assert np.array_equal(a, b), f"level {i} differs (seed {seed})"


def test_level_sizes_match_crate_b_dimension_test(oracle):
    # the only pyramid assertion in the reference: feature_tracker/src/image_operations.rs:84-94
    img = np.full((60, 120), 77, np.uint8)
    lv = oracle.split_pyramid(oracle.build_pyramid(img, 3), 120, 60, 3)
    assert [a.shape for a in lv] == [(60, 120), (30, 60), (15, 30)]


def test_level0_is_a_copy_and_constant_stays_constant(oracle):
    img = random_image(100, 50, 3)
    lv = oracle.split_pyramid(oracle.build_pyramid(img, 4), 100, 50, 4)
    assert np.array_equal(lv[0], img)
    for v in (0, 1, 128, 254, 255):
        c = np.full((50, 100), v, np.uint8)
        for a in oracle.split_pyramid(oracle.build_pyramid(c, 4), 100, 50, 4):
            assert (a == v).all()


def test_exact_2x_interior_weights(oracle):
    # exact 2x: interior taps are [1/8, 3/8, 3/8, 1/8] on window [2o-1, 2o+2] per axis
    img = random_image(64, 48, 4)
    half = oracle.resize_triangle(img, 32, 24)
    k = np.array([1, 3, 3, 1], np.float64) / 8.0
    f = img.astype(np.float64)
    for oy in range(1, 23):
        for ox in range(1, 31):
            win = f[2 * oy - 1:2 * oy + 3, 2 * ox - 1:2 * ox + 3]
            exact = k @ win @ k
            assert half[oy, ox] == int(np.floor(exact + 0.5)), (ox, oy)


def test_ties_round_half_away_from_zero(oracle):
    # 2x2 blocks of (0, 1) columns average to exactly x.5 in the interior -> must round up
    img = np.zeros((16, 16), np.uint8)
    img[:, 1::2] = 1
    half = oracle.resize_triangle(img, 8, 8)
    assert (half[1:-1, 1:-1] == 1).all()


def test_pillow_bilinear_agrees_within_one_grey_level(oracle):
    # Pillow's BILINEAR reduce uses the same tent support and centre convention but fixed-point
    # coefficients and a horizontal-first u8 intermediate: an independent sanity check of geometry.
    PIL = pytest.importorskip("PIL.Image")
    img = scene_pair(200, 120, 9)[0]
    for nw, nh in ((100, 60), (50, 30), (25, 15)):
        ours = oracle.resize_triangle(img, nw, nh).astype(int)
        theirs = np.asarray(PIL.fromarray(img).resize((nw, nh), PIL.BILINEAR)).astype(int)
        assert np.abs(ours - theirs).max() <= 1


def test_complete_dyadic_window_chain_is_an_exact_integer_tent():
    """The claim behind `vtent_exact4` (csrc/pyramid.cu): on an axis whose ratio is r = 2^lev, an output whose window
    is complete (2r taps) has the weights (2k+1)/(2r^2), mirrored, exactly in f32, and the reference's chain
    `t += pix * w` in tap order never rounds -- so t == N / (2 r^2) bit for bit, N the integer tent sum.  Checked on
    the numpy restatement of `image`'s sampler (oracle/np_crosscheck.py) for r = 16, 32, 64, 128 with worst-case
    (all 255) and random columns; clamped border windows (1.5 r taps) are NOT exact and keep their chains."""
    import math
    import numpy as np
    import sys, os
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
    import np_crosscheck as nc
    F = np.float32
    rng = np.random.default_rng(11)
    for r in (16, 32, 64, 128):
        n_out = 5
        n_in = r * n_out
        taps = nc._taps(n_in, n_out)
        ks = np.arange(2 * r)
        tent = np.where(ks < r, 2 * ks + 1, 4 * r - 2 * ks - 1).astype(np.int64)
        for o, (left, w) in enumerate(taps):
            if o in (0, n_out - 1):
                assert len(w) == r + r // 2            # clamped: renormalised by a non-power-of-two sum
                assert not np.array_equal(w * F(2 * r * r), np.round(w * F(2 * r * r)))
                continue
            assert left == o * r - r // 2 and len(w) == 2 * r
            assert np.array_equal(w.astype(np.float64) * (2 * r * r), tent.astype(np.float64))   # exact weights
            for col in (np.full(2 * r, 255, np.uint8), rng.integers(0, 256, 2 * r).astype(np.uint8),
                        (np.arange(2 * r) % 2 * 255).astype(np.uint8)):
                t = F(0)
                for k in range(2 * r):                 # the reference's chain, every operation rounded to f32
                    t = F(t + F(F(col[k]) * w[k]))
                n = int((tent * col.astype(np.int64)).sum())
                assert n < (1 << 24)
                assert t == F(n) * F(1.0 / (2 * r * r)) and float(t) * (2 * r * r) == n, (r, o, n, t)
