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
    rows_f()


def rows_f():
    """rows f-1 / f-3 (golden_rows_f.npz): unprojection with the cameras of the reference's configs and
    DynamicImage::to_luma8 of 16-bit / RGB frames, by the oracle."""
    out = {}
    rng = np.random.default_rng(20261018)
    euroc = [458.654, 457.296, 367.215, 248.375, -0.28340811, 0.07395907, 0.00019359, 1.76187114e-05, 0.0]
    tumvi = [191.75556798912652, 191.74816751185256, 254.9226487139376, 256.8780365577954, 0.6246288732884442,
             1.0598071085569876, 0.0, 0.0, 0.0]
    for name, model, P, (w, h) in (("euroc", orc.CAMERA_PINHOLE_RADTAN, euroc, (752, 480)),
                                   ("tumvi", orc.CAMERA_EUCM, tumvi, (512, 512))):
        uv = rng.uniform([0, 0], [w, h], (400, 2)).astype(np.float32)
        out[f"unproject_{name}_params"] = np.array(P)
        out[f"unproject_{name}_uv"] = uv
        out[f"unproject_{name}_xy"] = orc.unproject(model, P, uv.astype(np.float64))
    g16 = rng.integers(0, 65536, (33, 47), dtype=np.uint16)
    rgb = rng.integers(0, 256, (33, 47, 3), dtype=np.uint8)
    out["luma16_in"], out["luma16_out"] = g16, orc.to_luma8(g16, orc.PIXEL_LUMA16)
    out["rgb8_in"], out["rgb8_out"] = rgb, orc.to_luma8(rgb, orc.PIXEL_RGB8)
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden_rows_f.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    if "--rows-f" in sys.argv:   # only the second file (golden_small.npz stays byte-identical)
        rows_f()
    else:
        main()
