"""Writes a file in the format of tools/reference_vectors/src/main.rs, but with the ORACLE's answers.
It pins nothing -- it exists to exercise the file format and tests/test_reference_vectors.py without cargo:

    python tools/reference_vectors/emulate_with_oracle.py /tmp/vec.txt
    RSVIO_REFERENCE_VECTORS=/tmp/vec.txt python -m pytest tests/test_reference_vectors.py -q
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
from oracle import oracle  # noqa: E402
from test_reference_vectors import noise, scene  # noqa: E402


def bits(x):
    return "%08x" % int(np.array([x], np.float32).view(np.uint32)[0])


def main():
    out = ["# oracle-made stand-in for rsvio reference vectors v1 (pins nothing)"]
    for w, h, levels, seed, kind in ((64, 48, 4, 1, 0), (80, 60, 3, 2, 1), (107, 83, 5, 3, 1), (256, 192, 6, 4, 0), (752, 480, 6, 5, 0)):
        img = scene(w, h, 0, 0, seed) if kind == 0 else noise(w, h, seed)
        lv = oracle.split_pyramid(oracle.build_pyramid(img, levels), w, h, levels)
        for k in range(levels):
            out.append(f"pyr {w} {h} {levels} {seed} {kind} {k} {lv[k].size} : {lv[k].tobytes().hex()}")
    for w, h, seed, thr in ((40, 40, 11, 10), (40, 40, 11, 25), (40, 40, 11, 40), (64, 64, 12, 10)):
        xs, ys, sc = oracle.corners_fast9(scene(w, h, 3, 5, seed), thr)
        out.append(f"fast9 {w} {h} {seed} {thr} {len(xs)} : " + " ".join(f"{x},{y},{bits(s)}" for x, y, s in zip(xs, ys, sc)))
    for w, h, seed, grid in ((256, 192, 21, 30), (376, 240, 22, 40)):
        for occ in (0, 1):
            cur = [[40, 40], [w - 45, h - 50], [w // 2, h // 2]] if occ else np.zeros((0, 2), np.uint32)
            xs, ys, sc = oracle.detect_key_points(scene(w, h, 1, 2, seed), grid, cur)
            out.append(f"detect {w} {h} {seed} {grid} {occ} {len(xs)} : " + " ".join(f"{x},{y},{bits(s)}" for x, y, s in zip(xs, ys, sc)))
    for a in ([0.0, 0.0, 0.0], [1.5, -2.25, 0.0], [0.3, 0.7, 0.01], [-1.0, 2.0, 0.4], [0.02, -0.01, -1.2], [0.5, 0.5, 1e-8]):
        m = oracle.se2_exp(a)
        out.append(f"se2exp {bits(a[0])} {bits(a[1])} {bits(a[2])} : " + " ".join(bits(v) for v in m.ravel()))
    img = scene(96, 72, 2, 1, 31)
    for x, y in ((10.0, 10.0), (20.25, 31.75), (47.5, 35.5), (80.125, 60.875), (3.0, 3.0)):
        g = oracle.image_grad(img, x, y)
        out.append(f"grad 96 72 31 {bits(x)} {bits(y)} : {bits(g[0])} {bits(g[1])} {bits(g[2])}")
    for x, y in ((30.0, 30.0), (48.5, 36.25), (5.0, 5.0), (90.0, 66.0)):
        p = oracle.pattern52_new(img, x, y)
        out.append(f"pattern 96 72 31 {bits(x)} {bits(y)} {int(p.valid)} {bits(p.mean)} : " + " ".join(bits(v) for v in p.data[:]))
    w, h, seed = 256, 192, 41
    tr = oracle.StereoTracker(3, 30, 30, 0.005, nthreads=4)
    for t in range(5):
        tr.process_mono(scene(w, h, 2 * t, t, seed))
        xy = tr.get_tracks(0)[1][:, [2, 5]]
        out.append(f"track {w} {h} {seed} {t} {len(xy)} : " + " ".join(f"{bits(x)},{bits(y)}" for x, y in xy))
    from test_reference_vectors import EUROC_L, TUMVI_L

    def b64(x):
        return "%016x" % int(np.array([x], np.float64).view(np.uint64)[0])
    for model, P, pts in ((1, EUROC_L, ((367.215, 248.375), (10.0, 10.0), (740.0, 470.0), (100.5, 300.25))),
                          (2, TUMVI_L, ((254.9226487139376, 256.8780365577954), (100.5, 300.25), (400.0, 120.0)))):
        for u, v in pts:
            x, y = oracle.unproject(model, P, [[u, v]])[0]
            out.append(f"unproject {model} {b64(u)} {b64(v)} : {b64(x)} {b64(y)} {b64(1.0)}")
    from test_reference_vectors import mix
    yy, xx = np.mgrid[0:8, 0:64]
    g16 = (mix(xx, yy, 51) >> np.uint64(16)).astype(np.uint16)
    out.append(f"luma16 64 8 51 : {g16.astype('<u2').tobytes().hex()} {oracle.to_luma8(g16, oracle.PIXEL_LUMA16).tobytes().hex()}")
    m = mix(xx, yy, 52)
    rgb = np.stack([(m >> np.uint64(24)) & np.uint64(255), (m >> np.uint64(16)) & np.uint64(255), (m >> np.uint64(8)) & np.uint64(255)], 2).astype(np.uint8)
    out.append(f"rgb8 64 8 52 : {rgb.tobytes().hex()} {oracle.to_luma8(rgb, oracle.PIXEL_RGB8).tobytes().hex()}")
    path = sys.argv[1] if len(sys.argv) > 1 else "reference_vectors_oracle.txt"
    open(path, "w").write("\n".join(out) + "\n")
    print("wrote", path)


if __name__ == "__main__" and "--crate-b" not in sys.argv:
    main()


def crate_b(path):
    """same for tools/reference_vectors_b (crate B): `python emulate_with_oracle.py --crate-b /tmp/vec_b.txt`, then
    RSVIO_REFERENCE_VECTORS_B=/tmp/vec_b.txt python -m pytest tests/test_crate_b_oracle.py"""
    from oracle import crate_b as B
    from test_crate_b_oracle import scene_b

    def dump(v):
        return " ".join(bits(x) for x in np.asarray(v, np.float32).ravel())
    out = ["# oracle-made stand-in for the crate B vectors (pins nothing)"]
    for w, h, seed, levels, ratio in ((120, 60, 61, 3, 2.0), (97, 75, 62, 4, 2.0), (96, 64, 63, 3, 1.5)):
        pyr = B.build_image_pyramid(scene_b(w, h, seed), levels, ratio)
        for l, lv in enumerate(pyr):
            out.append(f"pyr_b {w} {h} {seed} {levels} {bits(ratio)} 0 {bits(0.0)} {l} {lv.shape[1]} {lv.shape[0]} : {dump(lv)}")
    img = scene_b(96, 72, 64)
    for x, y in ((10.0, 10.0), (20.25, 31.75), (47.5, 35.5), (80.125, 60.875), (0.5, 3.0), (93.2, 10.0)):
        r = B.d_interpolate_bicubic(img, x, y)
        if r is None:
            out.append(f"bicubic 96 72 64 {bits(x)} {bits(y)} 0 :")
        else:
            out.append(f"bicubic 96 72 64 {bits(x)} {bits(y)} 1 : {bits(r[0])} {bits(B.interpolate_bicubic(img, x, y))} {bits(r[1][0])} {bits(r[1][1])}")
    for x, y in ((30.0, 30.0), (48.5, 36.25), (4.0, 4.0)):
        p = B.Patch52(img, (x, y), 0.1)
        out.append(f"patch_b 96 72 64 {bits(x)} {bits(y)} : {dump(p.intensities)} | {dump(p.J)} | {dump(p.Hinv)}")
    open(path, "w").write("\n".join(out) + "\n")
    print("wrote", path)


if __name__ == "__main__" and "--crate-b" in sys.argv:
    crate_b(sys.argv[-1])
