"""GPU: grid FAST-9 replenishment (through rsvio_detect_keypoints) is identical to the oracle."""
import os

import numpy as np
import pytest

from util import GOLDEN_DIR, random_image, scene_pair

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("w,h,g", [(752, 480, 40), (752, 480, 50), (512, 512, 30), (800, 400, 12),
                                   (1920, 1200, 15), (203, 157, 9), (320, 240, 7)])
def test_detect_matches_oracle(rsvio, oracle, w, h, g):
    img = scene_pair(w, h, 7 + g)[0]
    rng = np.random.default_rng(g)
    cur = np.stack([rng.integers(0, w, 60), rng.integers(0, h, 60)], axis=1).astype(np.uint32)
    for occupied in (np.zeros((0, 2), np.uint32), cur):
        xs, ys, sc = oracle.detect_key_points(img, g, occupied)
        xy, s2 = rsvio.detect_keypoints(img, g, occupied)
        assert len(xy) == len(xs)
        assert np.array_equal(xy, np.stack([xs, ys], axis=1).astype(np.uint32))
        assert np.array_equal(s2, sc)


def test_detect_on_noise_and_flat(rsvio, oracle):
    img = random_image(400, 300, 3)
    xs, ys, sc = oracle.detect_key_points(img, 25, np.zeros((0, 2), np.uint32))
    xy, s2 = rsvio.detect_keypoints(img, 25)
    assert np.array_equal(xy, np.stack([xs, ys], axis=1).astype(np.uint32)) and np.array_equal(s2, sc)
    xy, _ = rsvio.detect_keypoints(np.full((300, 400), 90, np.uint8), 25)
    assert len(xy) == 0


def test_detect_golden(rsvio):
    g = np.load(os.path.join(GOLDEN_DIR, "golden_small.npz"))
    xy, sc = rsvio.detect_keypoints(g["det_img_a"], int(g["det_grid"]), g["det_cur"])
    assert np.array_equal(xy, g["det_xy"]) and np.array_equal(sc, g["det_score"])
