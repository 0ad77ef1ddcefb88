"""GPU: pyramid kernel (through the C ABI) is bit-exact against the CPU oracle."""
import os

import numpy as np
import pytest

from util import GOLDEN_DIR, random_image, scene_pair

pytestmark = pytest.mark.gpu

CASES = [
    ("C1", 752, 480, 3), ("EUROC_SHIPPED", 752, 480, 6), ("C2", 512, 512, 4), ("C3", 800, 400, 6),
    ("odd", 107, 83, 5), ("odd2", 333, 251, 6), ("tiny", 40, 33, 3), ("wide", 1000, 64, 4), ("l1", 64, 48, 1),
    ("deep", 1024, 768, 8),
]


@pytest.mark.parametrize("name,w,h,levels", CASES)
def test_pyramid_bit_exact(rsvio, oracle, name, w, h, levels):
    imgs = np.stack([random_image(w, h, 11), scene_pair(w, h, 12)[0], np.full((h, w), 255, np.uint8)])
    got = rsvio.build_pyramid(imgs, levels)
    for j in range(imgs.shape[0]):
        ref = oracle.build_pyramid(imgs[j], levels, nthreads=4)
        diff = np.flatnonzero(got[j] != ref)
        assert diff.size == 0, f"{name} image {j}: {diff.size} bytes differ, first at {diff[:5]}"


def test_pyramid_c4_full_size(rsvio, oracle):
    w, h, levels = 1920, 1200, 5
    img = scene_pair(w, h, 13)[0]
    got = rsvio.build_pyramid(img, levels)[0]
    assert np.array_equal(got, oracle.build_pyramid(img, levels, nthreads=8))
    lv = rsvio.split_pyramid(got, w, h, levels)
    assert np.array_equal(lv[0], img)  # level 0 is the input


def test_pyramid_golden(rsvio):
    g = np.load(os.path.join(GOLDEN_DIR, "golden_small.npz"))
    got = rsvio.build_pyramid(g["pyr_img"], int(g["pyr_levels"]))[0]
    assert np.array_equal(got, g["pyr_flat"])


def test_pyramid_constant_and_batch_independence(rsvio):
    w, h, levels = 752, 480, 3
    imgs = np.stack([np.full((h, w), v, np.uint8) for v in (0, 1, 127, 254, 255)])
    out = rsvio.build_pyramid(imgs, levels)
    for j, v in enumerate((0, 1, 127, 254, 255)):
        assert (out[j] == v).all()
    # a batch gives the same bytes as single-image calls
    rnd = np.stack([random_image(w, h, 100 + j) for j in range(4)])
    batch = rsvio.build_pyramid(rnd, levels)
    for j in range(4):
        assert np.array_equal(batch[j], rsvio.build_pyramid(rnd[j], levels)[0])
