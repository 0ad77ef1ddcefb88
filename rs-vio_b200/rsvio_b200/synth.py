"""Synthetic stereo sequences with exact sub-pixel ground truth.

A scene is an analytic texture
    I(x, y) = clamp(128 + sum_k A_k sin(2*pi*(u_k x + v_k y) + phi_k) + sum_j c_j * cover_j(x, y), 0, 255)
with 24 sinusoids (amplitude 4..20, wavelength 6..80 px) and box-filtered axis-aligned
rectangles (8..40 px, contrast +-[50,100], about one per 30x30 px, so every grid cell sees
FAST-9 corners).  Frame t of the left camera samples I(x - t*dx, y - t*dy); the right camera
samples the left image shifted by the disparity.  A point seen at p in the left frame 0 is
therefore at p + t*(dx,dy) in left frame t and at p + t*(dx,dy) + disparity in the right one.
Everything is seeded (numpy PCG64), so tests, bench and the CPU baseline see the same bytes.
"""
from dataclasses import dataclass

import numpy as np

# The five BASELINE.json configs (SURVEY.md section 8 shapes) + the estimator's shipped setting.
CONFIGS = {
    "C1": dict(w=752, h=480, levels=3, grid=40, max_it=20, thr=0.01, motion=(1.37, -0.62), disparity=(-8.25, 0.0)),
    "C2": dict(w=512, h=512, levels=4, grid=30, max_it=10, thr=0.01, motion=(1.37, -0.62), disparity=(-8.25, 0.0)),
    "C3": dict(w=800, h=400, levels=6, grid=12, max_it=10, thr=0.01, motion=(3.25, 0.4), disparity=(-8.25, 0.0),
               dyn_frac=0.15),
    "C4": dict(w=1920, h=1200, levels=5, grid=15, max_it=20, thr=0.01, motion=(1.37, -0.62),
               disparity=(-20.5, 0.0)),
    "C5": dict(w=752, h=480, levels=3, grid=40, max_it=20, thr=0.01, motion=(1.37, -0.62), disparity=(-8.25, 0.0),
               streams=64),
    "EUROC_SHIPPED": dict(w=752, h=480, levels=6, grid=50, max_it=20, thr=0.01, motion=(1.37, -0.62),
                          disparity=(-8.25, 0.0)),
}
CONFIG_INDEX = {"C1": 1, "C2": 2, "C3": 3, "C4": 4, "C5": 5, "EUROC_SHIPPED": 6}


def stream_seed(config, stream_id=0):
    """seed = 42 + 1000*config + stream (42 echoes reference src/bin/run_euroc.rs:11)."""
    return 42 + 1000 * CONFIG_INDEX[config] + stream_id


@dataclass
class _Rect:
    x0: float
    y0: float
    x1: float
    y1: float
    c: float


class Scene:
    """One seeded analytic texture; render(shift) samples I(x - sx, y - sy) on the pixel grid."""

    def __init__(self, w, h, seed, n_sin=24, rect_pitch=30.0, margin=160):
        self.w, self.h = int(w), int(h)
        rng = np.random.default_rng(seed)
        lam = rng.uniform(6.0, 80.0, n_sin)
        ang = rng.uniform(0.0, 2 * np.pi, n_sin)
        self.u = (np.cos(ang) / lam)
        self.v = (np.sin(ang) / lam)
        self.amp = rng.uniform(4.0, 20.0, n_sin)
        self.phi = rng.uniform(0.0, 2 * np.pi, n_sin)
        area = (w + 2 * margin) * (h + 2 * margin)
        n_rect = int(area / (rect_pitch * rect_pitch))
        x0 = rng.uniform(-margin, w + margin, n_rect)
        y0 = rng.uniform(-margin, h + margin, n_rect)
        ww = rng.uniform(8.0, 40.0, n_rect)
        hh = rng.uniform(8.0, 40.0, n_rect)
        cc = rng.uniform(50.0, 100.0, n_rect) * rng.choice([-1.0, 1.0], n_rect)
        self.rects = [_Rect(a, b, a + c, b + d, e) for a, b, c, d, e in zip(x0, y0, ww, hh, cc)]
        self._dyn_seed = int(seed) * 7919 + 13

    def _field(self, shift):
        sx, sy = float(shift[0]), float(shift[1])
        xs = np.arange(self.w, dtype=np.float64) - sx
        ys = np.arange(self.h, dtype=np.float64) - sy
        # sin(a_x + b_y + phi) = sin(b_y + phi) cos(a_x) + cos(b_y + phi) sin(a_x): one matmul
        ax = 2 * np.pi * np.outer(self.u, xs)                      # K x w
        by = 2 * np.pi * np.outer(ys, self.v) + self.phi[None, :]  # h x K
        left = np.concatenate([np.sin(by) * self.amp[None, :], np.cos(by) * self.amp[None, :]], axis=1)
        right = np.concatenate([np.cos(ax), np.sin(ax)], axis=0)
        img = 128.0 + left @ right
        for r in self.rects:
            # exact area coverage of pixel [x-.5, x+.5) by the rectangle, separable
            xa, xb = r.x0 + sx, r.x1 + sx
            ya, yb = r.y0 + sy, r.y1 + sy
            ix0, ix1 = max(0, int(np.floor(xa - 0.5))), min(self.w, int(np.ceil(xb + 0.5)) + 1)
            iy0, iy1 = max(0, int(np.floor(ya - 0.5))), min(self.h, int(np.ceil(yb + 0.5)) + 1)
            if ix0 >= ix1 or iy0 >= iy1:
                continue
            px = np.arange(ix0, ix1, dtype=np.float64)
            py = np.arange(iy0, iy1, dtype=np.float64)
            cx = np.clip(np.minimum(px + 0.5, xb) - np.maximum(px - 0.5, xa), 0.0, 1.0)
            cy = np.clip(np.minimum(py + 0.5, yb) - np.maximum(py - 0.5, ya), 0.0, 1.0)
            img[iy0:iy1, ix0:ix1] += r.c * np.outer(cy, cx)
        return img

    def render(self, shift=(0.0, 0.0), dyn_frac=0.0, dyn_key=0):
        """u8 image of I(x - shift).  dyn_frac > 0 overwrites a vertical band (that fraction of the
        width, fixed in image coordinates) with texture that depends on dyn_key, so points inside it
        cannot be tracked consistently (exercises forward-backward rejection)."""
        img = self._field(shift)
        if dyn_frac > 0.0:
            bw = int(self.w * dyn_frac)
            b0 = (self.w - bw) // 2
            rng = np.random.default_rng(self._dyn_seed + 104729 * int(dyn_key))
            coarse = rng.uniform(0.0, 255.0, (self.h // 4 + 2, bw // 4 + 2))
            band = np.kron(coarse, np.ones((4, 4)))[: self.h, :bw]
            img[:, b0:b0 + bw] = band
        img = np.clip(img, 0.0, 255.0)
        return np.floor(img + 0.5).astype(np.uint8)  # round half away from zero (values >= 0)


class StereoSequence:
    """Frames of one synthetic stereo stream of a named config."""

    def __init__(self, config="C1", stream_id=0, **overrides):
        p = dict(CONFIGS[config])
        p.update(overrides)
        self.params = p
        self.config = config
        self.scene = Scene(p["w"], p["h"], stream_seed(config, stream_id))

    def frame(self, t):
        p = self.params
        mx, my = p["motion"]
        dx, dy = p["disparity"]
        dyn = p.get("dyn_frac", 0.0)
        left = self.scene.render((t * mx, t * my), dyn, 2 * t)
        right = self.scene.render((t * mx + dx, t * my + dy), dyn, 2 * t + 1)
        return left, right

    def true_left(self, p0, t):
        mx, my = self.params["motion"]
        return np.asarray(p0, np.float64) + t * np.array([mx, my])

    def true_right(self, p_left):
        return np.asarray(p_left, np.float64) + np.array(self.params["disparity"])
