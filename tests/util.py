"""Shared helpers for the parity tests (seeded synthetic inputs, golden paths)."""
import os

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def scene_pair(w, h, seed, shift=(1.37, -0.62)):
    """two frames of one seeded analytic scene: I(x, y) and I(x - sx, y - sy)"""
    from rsvio_b200.synth import Scene
    sc = Scene(w, h, seed)
    return sc.render((0.0, 0.0)), sc.render(shift)


def random_image(w, h, seed):
    rng = np.random.default_rng(seed)
    return rng.integers(0, 256, (h, w), dtype=np.uint8)


def grid_points(w, h, n, seed, margin=24.0):
    rng = np.random.default_rng(seed)
    xy = np.stack([rng.uniform(margin, w - margin, n), rng.uniform(margin, h - margin, n)], axis=1)
    return xy.astype(np.float32)


def identity_T6(xy):
    xy = np.asarray(xy, np.float32).reshape(-1, 2)
    T = np.zeros((xy.shape[0], 6), np.float32)
    T[:, 0] = 1.0
    T[:, 4] = 1.0
    T[:, 2] = xy[:, 0]
    T[:, 5] = xy[:, 1]
    return T
