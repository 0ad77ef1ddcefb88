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


def test_streaming_kernel_variant_is_bit_exact_too():
    """RSVIO_PYR_KERNEL=stream selects the experimental warp-per-strip kernel (read once per process):
    run the same oracle comparison in a child process."""
    import subprocess
    import sys
    code = (
        "import sys, numpy as np\n"
        "sys.path.insert(0, %r); sys.path.insert(0, %r); sys.path.insert(0, %r)\n"
        "import rsvio_b200\n"
        "from oracle import oracle as orc\n"
        "from util import random_image, scene_pair\n"
        "for (w, h, lv) in ((752, 480, 3), (512, 512, 4), (1920, 1200, 5), (752, 480, 6)):\n"
        "    imgs = np.stack([random_image(w, h, 5), scene_pair(w, h, 6)[0]])\n"
        "    got = rsvio_b200.build_pyramid(imgs, lv)\n"
        "    for j in range(2):\n"
        "        assert np.array_equal(got[j], orc.build_pyramid(imgs[j], lv, nthreads=4)), (w, h, lv, j)\n"
        "print('stream ok')\n"
    ) % (os.path.join(os.path.dirname(GOLDEN_DIR), ".."), os.path.join(os.path.dirname(GOLDEN_DIR), "..", "rs-vio_b200"),
         os.path.dirname(GOLDEN_DIR))
    env = dict(os.environ, RSVIO_PYR_KERNEL="stream")
    out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and "stream ok" in out.stdout, out.stdout + out.stderr


def test_pyramid_random_shapes_fuzz(rsvio, oracle):
    """random sizes (tile-aligned, 16-aligned, arbitrary), level counts, batch sizes and tie-heavy content:
    every byte equals the oracle's (the tile hand-out, the border ring and the generic tiling all depend on shape)"""
    rng = np.random.default_rng(2024)
    for c in range(36):
        kind = c % 4
        if kind == 0:
            m = int(rng.choice([16, 32, 64]))
            w, h = m * int(rng.integers(2, 30)), m * int(rng.integers(1, 16))
        elif kind == 1:
            w, h = 16 * int(rng.integers(2, 50)), int(rng.integers(20, 500))
        elif kind == 2:
            w, h = int(rng.integers(20, 900)), int(rng.integers(20, 600))
        else:
            w = 256 * int(rng.integers(1, 4)) + 16 * int(rng.integers(0, 3))
            h = 64 * int(rng.integers(1, 6)) + int(rng.choice([0, 16, 32]))
        levels = int(rng.integers(1, 8))
        while levels > 1 and ((w >> (levels - 1)) < 1 or (h >> (levels - 1)) < 1):
            levels -= 1
        n = int(rng.integers(1, 4))
        if c % 3 == 0:
            img = rng.integers(0, 256, (n, h, w), dtype=np.uint8)
        elif c % 3 == 1:   # few grey values: many exact x.5 ties
            img = (rng.integers(0, 4, (n, h, w)) * 64 + rng.integers(0, 2, (n, h, w))).astype(np.uint8)
        else:
            img = np.full((n, h, w), int(rng.integers(0, 256)), np.uint8)
            img[:, ::7, ::5] = 255
        got = rsvio.build_pyramid(img, levels)
        for j in range(n):
            want = oracle.build_pyramid(img[j], levels, nthreads=4)
            assert np.array_equal(got[j], want), f"{w}x{h} L{levels} n={n} case {c} image {j}"
