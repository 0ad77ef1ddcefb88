"""Pins against third-party implementations that ship in the image (OpenCV 4.13, Pillow) -- see thirdparty.py.

The reference delegates three pieces of hot-path arithmetic to crates that are not in its tree (`image` resize,
`imageproc` corners_fast9, `camera-intrinsic-model` unprojection) and holds no test for them, and cargo is not in
this image.  These tests check the oracle (CPU, default run) and the CUDA path (-m gpu, through the C ABI) against
code neither of them shares a line with:

  pinned here                                        by
  FAST-9 corner SETS at thresholds 10..40            cv2.FastFeatureDetector (TYPE_9_16, no NMS)
  FAST-9 scores, per-cell pick of detect_key_points  the same corner sets (score = largest threshold still a corner)
  radtan unprojection (row f-1)                      cv2.undistortPointsIter
  pyramid levels 1..3, complete windows              integer tent sum, round half up
  pyramid levels 1..5 incl. borders and 752->23      Pillow BILINEAR reduce, +-1 grey level

What stays unpinned after them is listed in DESIGN.md section 3.
"""
import numpy as np
import pytest

import thirdparty as tp
from test_unproject import EUROC_L, EUROC_R, grid

cv2 = pytest.importorskip("cv2")


# ---- FAST-9 ------------------------------------------------------------------------------------------
def test_oracle_fast9_sets_and_scores_equal_opencv(oracle):
    for i, img in enumerate(tp.pin_images(240, 180)):
        sets = tp.cv_fast9_sets(img, 10, 40)
        for t in range(10, 41):
            xs, ys, _ = oracle.corners_fast9(img, t)
            m = np.zeros(img.shape, bool)
            m[ys, xs] = True
            want = sets[t] if t in sets else np.zeros(img.shape, bool)
            assert np.array_equal(m, want), (i, t)
        smap = tp.cv_fast9_score_map(img, 10)
        xs, ys, sc = oracle.corners_fast9(img, 10)
        assert (smap >= 10).sum() == len(xs)
        assert np.array_equal(smap[ys, xs], sc.astype(np.int32)), i


DETECT_CASES = [(752, 480, 40, (0, 2)), (512, 512, 30, (1, 3)), (203, 157, 15, (0, 1, 2, 3, 4)), (320, 240, 7, (2, 4))]


def _occupied(w, h, g):
    rng = np.random.default_rng(g)
    return np.stack([rng.integers(0, w, 60), rng.integers(0, h, 60)], axis=1).astype(np.uint32)


@pytest.mark.parametrize("w,h,g,which", DETECT_CASES)
def test_oracle_detect_key_points_equals_opencv_driven_pick(oracle, w, h, g, which):
    imgs = tp.pin_images(w, h)
    for k in which:
        for occ in (np.zeros((0, 2), np.uint32), _occupied(w, h, g)):
            xs, ys, sc = oracle.detect_key_points(imgs[k], g, occ)
            got = [(int(a), int(b), float(c)) for a, b, c in zip(xs, ys, sc)]
            assert got == tp.cv_detect_key_points(imgs[k], g, occ), (w, h, g, k)


@pytest.mark.gpu
@pytest.mark.parametrize("w,h,g,which", DETECT_CASES + [(1920, 1200, 15, (0,))])
def test_gpu_detect_key_points_equals_opencv_driven_pick(rsvio, w, h, g, which):
    imgs = tp.pin_images(w, h)
    for k in which:
        for occ in (np.zeros((0, 2), np.uint32), _occupied(w, h, g)):
            xy, sc = rsvio.detect_keypoints(imgs[k], g, occ)
            got = [(int(a), int(b), float(c)) for (a, b), c in zip(xy, sc)]
            assert got == tp.cv_detect_key_points(imgs[k], g, occ), (w, h, g, k)


# ---- unprojection (row f-1) ----------------------------------------------------------------------------
def _cv_undistort(K, D, uv):
    Kmat = np.array([[K[0], 0, K[2]], [0, K[1], K[3]], [0, 0, 1]], np.float64)
    dist = np.array((list(D) + [0.0] * 5)[:5], np.float64)          # k1 k2 p1 p2 k3
    crit = (cv2.TERM_CRITERIA_COUNT | cv2.TERM_CRITERIA_EPS, 200, 1e-14)
    return cv2.undistortPointsIter(uv.reshape(-1, 1, 2).astype(np.float64), Kmat, dist, None, None, crit).reshape(-1, 2)


def test_oracle_unproject_equals_opencv(oracle):
    for K, D in (EUROC_L, EUROC_R):
        uv = grid(752, 480, 7)
        want = _cv_undistort(K, D, uv)
        got = oracle.unproject(1, K + D + [0.0], uv)
        assert np.isfinite(got).all() and np.abs(got - want).max() < 1e-12


@pytest.mark.gpu
def test_gpu_unproject_equals_opencv(rsvio):
    for K, D in (EUROC_L, EUROC_R):
        uv = grid(752, 480, 7).astype(np.float32)
        want = _cv_undistort(K, D, uv)
        got = rsvio.unproject_points(rsvio.CameraModel.pinhole_radtan(K, D), uv)
        assert np.abs(got.astype(np.float64) - want).max() <= 1e-6


# ---- pyramid ----------------------------------------------------------------------------------------
PYR_CASES = [(752, 480, 6), (800, 400, 6), (512, 512, 4), (1920, 1200, 5), (203, 157, 4)]


def _check_pyramid(levels_u8, img, w, h, L):
    for i in range(1, min(L, 4)):
        if w % (1 << i) == 0 and h % (1 << i) == 0:
            assert np.array_equal(levels_u8[i][1:-1, 1:-1], tp.tent_level_int(img, i)), f"level {i}: integer tent"
    for i, ref in enumerate(tp.pillow_levels(img, L)):
        d = np.abs(levels_u8[i].astype(np.int32) - ref.astype(np.int32))
        assert d.max() <= (0 if i == 0 else 1), f"level {i}: Pillow differs by {d.max()}"


@pytest.mark.parametrize("w,h,L", PYR_CASES)
def test_oracle_pyramid_against_integer_tent_and_pillow(oracle, w, h, L):
    for img in tp.pin_images(w, h)[:2 if w > 1000 else 3]:
        _check_pyramid(oracle.split_pyramid(oracle.build_pyramid(img, L), w, h, L), img, w, h, L)


@pytest.mark.gpu
@pytest.mark.parametrize("w,h,L", PYR_CASES)
def test_gpu_pyramid_against_integer_tent_and_pillow(rsvio, w, h, L):
    for img in tp.pin_images(w, h)[:2 if w > 1000 else 3]:
        _check_pyramid(rsvio.split_pyramid(rsvio.build_pyramid(img, L)[0], w, h, L), img, w, h, L)
