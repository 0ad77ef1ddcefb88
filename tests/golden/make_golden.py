"""Generates tests/golden/golden_small.npz with the CPU oracle (oracle/rsvio_oracle.cpp).

The reference ships no fixtures for tracker A and cannot be run here (no Rust toolchain), so these
vectors pin the ORACLE's output, not the reference binary's ("parity unpinned").  They
guard the oracle against regressions and give the GPU path a second, frozen target.
Run from the repo root:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "rs-vio_b200"))

from oracle import oracle as orc  # noqa: E402
from rsvio_b200.synth import Scene, StereoSequence  # noqa: E402


def main():
    out = {}
    # pyramid: odd size so that deeper levels use non-power-of-two ratios and truncated windows
    w, h, levels = 107, 83, 5
    sc = Scene(w, h, 20261018)
    img = sc.render((0.0, 0.0))
    out["pyr_img"] = img
    out["pyr_levels"] = np.int32(levels)
    out["pyr_flat"] = orc.build_pyramid(img, levels)
    # detection
    w2, h2, g = 160, 120, 16
    sc2 = Scene(w2, h2, 777)
    a = sc2.render((0.0, 0.0))
    b = sc2.render((0.83, -1.21))
    cur = np.array([[30, 30], [100, 64], [159, 119]], np.uint32)
    xs, ys, scs = orc.detect_key_points(a, g, cur)
    out["det_img_a"], out["det_img_b"] = a, b
    out["det_grid"], out["det_cur"] = np.int32(g), cur
    out["det_xy"] = np.stack([xs, ys], axis=1).astype(np.uint32)
    out["det_score"] = scs
    # LK: detected corners of a tracked into b, 3 levels
    lv = 3
    pa, pb = orc.build_pyramid(a, lv), orc.build_pyramid(b, lv)
    T0 = orc.identity_transforms(out["det_xy"])
    T1, status, iters, margins = orc.track_points(pa, pb, w2, h2, lv, T0, 20, 0.01)
    out["lk_levels"] = np.int32(lv)
    out["lk_T0"], out["lk_T1"], out["lk_status"], out["lk_iters"], out["lk_margins"] = T0, T1, status, iters, margins
    # stateful stereo tracker, 3 frames at quarter EuRoC size
    seq = StereoSequence("C1", w=188, h=120, grid=20)
    tr = orc.StereoTracker(3, 20, 20, 0.01)
    for t in range(3):
        l, r = seq.frame(t)
        tr.process_frame(l, r)
        for cam in (0, 1):
            ids, T = tr.get_tracks(cam)
            out[f"seq_ids_{t}_{cam}"] = ids
            out[f"seq_T_{t}_{cam}"] = T
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden_small.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
