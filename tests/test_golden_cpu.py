"""CPU: the oracle still reproduces the committed golden vectors (tests/golden/make_golden.py)."""
import os

import numpy as np

from util import GOLDEN_DIR


def test_oracle_reproduces_golden(oracle):
    g = np.load(os.path.join(GOLDEN_DIR, "golden_small.npz"))
    img = g["pyr_img"]
    assert np.array_equal(oracle.build_pyramid(img, int(g["pyr_levels"])), g["pyr_flat"])
    a, b, grid = g["det_img_a"], g["det_img_b"], int(g["det_grid"])
    xs, ys, sc = oracle.detect_key_points(a, grid, g["det_cur"])
    assert np.array_equal(np.stack([xs, ys], axis=1), g["det_xy"]) and np.array_equal(sc, g["det_score"])
    lv = int(g["lk_levels"])
    h, w = a.shape
    pa, pb = oracle.build_pyramid(a, lv), oracle.build_pyramid(b, lv)
    T1, status, iters, _ = oracle.track_points(pa, pb, w, h, lv, g["lk_T0"], 20, 0.01)
    assert np.array_equal(status, g["lk_status"]) and np.array_equal(iters, g["lk_iters"])
    assert np.array_equal(T1, g["lk_T1"])


def test_oracle_reproduces_golden_rows_f(oracle):
    """rows f-1 / f-3: unprojection and to_luma8 (tests/golden/make_golden.py --rows-f)"""
    g = np.load(os.path.join(GOLDEN_DIR, "golden_rows_f.npz"))
    for name, model in (("euroc", oracle.CAMERA_PINHOLE_RADTAN), ("tumvi", oracle.CAMERA_EUCM)):
        got = oracle.unproject(model, list(g[f"unproject_{name}_params"]), g[f"unproject_{name}_uv"].astype(np.float64))
        assert np.array_equal(got, g[f"unproject_{name}_xy"], equal_nan=True)
    assert np.array_equal(oracle.to_luma8(g["luma16_in"], oracle.PIXEL_LUMA16), g["luma16_out"])
    assert np.array_equal(oracle.to_luma8(g["rgb8_in"], oracle.PIXEL_RGB8), g["rgb8_out"])
