"""LK parity statistics against the oracle without assertions: distribution of position differences, flag differences and
how many of them the oracle's own threshold margins explain at several edge widths.
Usage: python tools/lk_parity_probe.py   (RSVIO_LIB selects the library build)"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "rs-vio_b200"), os.path.join(ROOT, "tests")]
import numpy as np  # noqa: E402
import rsvio_b200 as rsvio  # noqa: E402
from oracle import oracle  # noqa: E402
from util import grid_points, identity_T6, scene_pair  # noqa: E402

CASES = [("corners C1", 752, 480, 3, 20, 40), ("corners C3-L6", 800, 400, 6, 10, 12), ("corners C4", 1920, 1200, 5, 20, 15),
         ("random 640x480 L4", 640, 480, 4, 15, 0)]
for tag, w, h, levels, max_it, grid in CASES:
    if grid:
        a, b = scene_pair(w, h, 1000 + levels, (1.37, -0.62))
        xs, ys, _ = oracle.detect_key_points(a, grid, np.zeros((0, 2), np.uint32))
        T0 = identity_T6(np.stack([xs, ys], axis=1))
    else:
        a, b = scene_pair(w, h, 77, (-2.4, 1.9))
        T0 = identity_T6(grid_points(w, h, 3000, 5, margin=-4.0))
    pa, pb = oracle.build_pyramid(a, levels, nthreads=4), oracle.build_pyramid(b, levels, nthreads=4)
    T_ref, st_ref, it_ref, mg = oracle.track_points(pa, pb, w, h, levels, T0, max_it, 0.01, nthreads=8)
    T_gpu, st_gpu, it_gpu = rsvio.track_points(pa, pb, w, h, levels, T0, max_it, 0.01)
    both = (st_ref == 0) & (st_gpu == 0)
    err = np.abs(T_gpu - T_ref)[:, [2, 5]].max(axis=1)[both]
    flag = (st_ref == 0) != (st_gpu == 0)
    same_it = (it_gpu == it_ref).all(axis=1)[both]
    pct = np.percentile(err, [50, 90, 99, 99.9]) if len(err) else [0] * 4
    line = f"{tag}: n={len(st_ref)} ok={both.sum()} flagdiff={flag.sum()} same_iters={same_it.mean():.4f} " \
           f"err p50={pct[0]:.1e} p90={pct[1]:.1e} p99={pct[2]:.1e} p99.9={pct[3]:.1e} max={err.max():.1e} >=1e-3: {(err >= 1e-3).sum()}"
    big = np.zeros(len(st_ref), bool)
    big[np.flatnonzero(both)[err >= 1e-3]] = True
    big |= flag
    for edge in (1e-5, 3e-5, 1e-4):
        ex = (mg[:, 0] < edge) | (mg[:, 1] < 1e-4) | (mg[:, 2] < 1e-4)
        line += f" | edge {edge:.0e}: excluded {ex.sum()} unexplained {(big & ~ex).sum()}"
    print(line, flush=True)
