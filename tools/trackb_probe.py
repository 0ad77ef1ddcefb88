"""Row f-2 first slice: wall time of rsvio_b_track_points (uploads + kernel + download) for many points on an
EuRoC-size 3-level f32 pyramid pair.  Usage: python tools/trackb_probe.py [n_points]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "rs-vio_b200"))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import rsvio_b200  # noqa: E402


def scene(h, w, dx=0.0, dy=0.0):
    y, x = np.mgrid[0:h, 0:w].astype(np.float64)
    x, y = x - dx, y - dy
    return (0.5 + 0.18 * np.sin(0.21 * x + 0.13 * y) + 0.14 * np.sin(0.09 * x - 0.17 * y + 1.0) + 0.1 * np.cos(0.3 * y)).astype(np.float32)


def pyr(img, levels):   # plain 2x2 box pyramid: the probe measures the tracker, not crate B's resize
    out = [img]
    for _ in range(1, levels):
        a = out[-1]
        h, w = a.shape[0] // 2 * 2, a.shape[1] // 2 * 2
        out.append(a[:h, :w].reshape(h // 2, 2, w // 2, 2).mean(axis=(1, 3)).astype(np.float32))
    return out


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
    h, w = 480, 752
    pa, pb = pyr(scene(h, w), 3), pyr(scene(h, w, 1.3, -0.7), 3)
    pts = np.random.default_rng(0).uniform([30, 30], [w - 30, h - 30], (n, 2)).astype(np.float32)
    rsvio_b200.b_track_points(pa, pb, pts[:64], 10, 0.1)
    t0 = time.perf_counter()
    T, st = rsvio_b200.b_track_points(pa, pb, pts, 10, 0.1)
    dt = time.perf_counter() - t0
    flow = np.stack([T[:, 0] * pts[:, 0] - T[:, 1] * pts[:, 1] + T[:, 2], T[:, 1] * pts[:, 0] + T[:, 0] * pts[:, 1] + T[:, 3]], 1) - pts
    ok = st == 0
    print(f"{n} points: {dt * 1e3:.2f} ms wall incl. pyramid upload ({n / dt / 1e6:.2f} M tracks/s), {ok.mean():.3f} tracked, "
          f"median flow {np.median(flow[ok], axis=0)}")


if __name__ == "__main__":
    main()
