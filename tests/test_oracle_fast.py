"""CPU: FAST-9 / grid replenishment of the oracle (literal binary-search score, literal sort)
against the closed-form numpy restatement.  PARITY UNPINNED."""
import numpy as np
import pytest

from oracle import np_crosscheck as npx
from util import random_image, scene_pair


@pytest.mark.parametrize("seed", [0, 1, 2])
@pytest.mark.parametrize("thr", [10, 25, 40])
def test_corners_fast9_matches_closed_form(oracle, seed, thr):
    img = scene_pair(120, 90, 100 + seed)[0] if seed else random_image(120, 90, 5)
    xs, ys, sc = oracle.corners_fast9(img, thr)
    x2, y2, s2 = npx.corners_fast9(img, thr)
    assert np.array_equal(xs, x2) and np.array_equal(ys, y2) and np.array_equal(sc, s2)
    assert (sc >= thr).all()


@pytest.mark.parametrize("w,h,g", [(752, 480, 40), (752, 480, 50), (512, 512, 30), (200, 160, 12), (203, 157, 15)])
def test_detect_key_points_matches_numpy(oracle, w, h, g):
    img = scene_pair(w, h, 7)[0]
    rng = np.random.default_rng(3)
    cur = np.stack([rng.integers(0, w, 40), rng.integers(0, h, 40)], axis=1).astype(np.uint32)
    for occupied in (np.zeros((0, 2), np.uint32), cur):
        xs, ys, sc = oracle.detect_key_points(img, g, occupied)
        ref = npx.detect_key_points(img, g, occupied)
        assert len(xs) == len(ref)
        assert [(int(a), int(b), float(c)) for a, b, c in zip(xs, ys, sc)] == ref


def test_detect_respects_border_and_one_per_cell(oracle):
    w, h, g = 320, 240, 20
    img = scene_pair(w, h, 11)[0]
    xs, ys, _ = oracle.detect_key_points(img, g, np.zeros((0, 2), np.uint32))
    assert len(xs) > 0
    assert (xs >= 19).all() and (xs <= w - 19).all() and (ys >= 19).all() and (ys <= h - 19).all()
    x0, y0 = (w % g) // 2, (h % g) // 2
    cells = set()
    for x, y in zip(xs.tolist(), ys.tolist()):
        c = ((x - x0) // g, (y - y0) // g)
        assert c not in cells
        cells.add(c)
    # visiting order: x outer, y inner
    order = [((x - x0) // g) * 10000 + (y - y0) // g for x, y in zip(xs.tolist(), ys.tolist())]
    assert order == sorted(order)


def test_occupied_cells_get_no_point(oracle):
    w, h, g = 320, 240, 40
    img = scene_pair(w, h, 12)[0]
    xs, ys, _ = oracle.detect_key_points(img, g, np.zeros((0, 2), np.uint32))
    cur = np.stack([xs, ys], axis=1).astype(np.uint32)
    x2, _, _ = oracle.detect_key_points(img, g, cur)
    assert len(x2) == 0
