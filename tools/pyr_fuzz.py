"""Random-shape fuzz of the pyramid path against the oracle (bit-exact or it prints the first difference).
Usage: python tools/pyr_fuzz.py [cases] [seed]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "rs-vio_b200"))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import rsvio_b200  # noqa: E402
from oracle import oracle  # noqa: E402  (test infrastructure: the checker)


def main():
    cases = int(sys.argv[1]) if len(sys.argv) > 1 else 40
    rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
    bad = 0
    for c in range(cases):
        kind = rng.integers(0, 4)
        if kind == 0:      # fast path: multiples of 16 / 32 / 64
            m = int(rng.choice([16, 32, 64]))
            w, h = m * int(rng.integers(2, 40)), m * int(rng.integers(1, 24))
        elif kind == 1:    # width multiple of 16, odd height
            w, h = 16 * int(rng.integers(2, 60)), int(rng.integers(20, 700))
        elif kind == 2:    # arbitrary
            w, h = int(rng.integers(20, 1300)), int(rng.integers(20, 900))
        else:              # tile-boundary widths
            w, h = 256 * int(rng.integers(1, 5)) + 16 * int(rng.integers(0, 3)), 64 * int(rng.integers(1, 8)) + int(rng.choice([0, 16, 32]))
        levels = int(rng.integers(1, 8))
        while levels > 1 and ((w >> (levels - 1)) < 1 or (h >> (levels - 1)) < 1):
            levels -= 1
        n = int(rng.integers(1, 4))
        mode = rng.integers(0, 3)
        if mode == 0:
            img = rng.integers(0, 256, (n, h, w), dtype=np.uint8)
        elif mode == 1:   # smooth + ties: many x.5 values
            img = (rng.integers(0, 4, (n, h, w)) * 64 + rng.integers(0, 2, (n, h, w))).astype(np.uint8)
        else:
            img = np.full((n, h, w), int(rng.integers(0, 256)), np.uint8)
            img[:, ::7, ::5] = 255
        got = rsvio_b200.build_pyramid(img, levels)          # (n, dense pyramid bytes)
        ok = True
        for j in range(n):
            want = oracle.build_pyramid(img[j], levels, nthreads=8)
            if not np.array_equal(got[j], want):
                gl, wl = rsvio_b200.split_pyramid(got[j], w, h, levels), rsvio_b200.split_pyramid(want, w, h, levels)
                for lev in range(levels):
                    d = np.argwhere(gl[lev] != wl[lev])
                    if len(d):
                        print(f"MISMATCH case {c}: {w}x{h} L{levels} n={n} mode={mode} image {j} level {lev}: "
                              f"{len(d)} bytes, first (y, x) = {d[0].tolist()}")
                        break
                ok = False
                break
        bad += 0 if ok else 1
    print(f"pyr_fuzz: {cases - bad} / {cases} shapes bit-exact")
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
