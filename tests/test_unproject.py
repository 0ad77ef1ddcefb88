"""Row f-1: unprojection of tracked pixels (Feature::undistorted_coord, estimator/frame.rs:107-134).
CPU: the oracle's models (round trips, known points).  GPU (-m gpu): the CUDA path against the oracle --
Newton on the device vs fixed-point iteration in the oracle, both in f64 -- and the fused export of a tracker.
Parity with the reference's crate (camera-intrinsic-model 0.7.2) is unpinned; see test_reference_vectors.py."""
import numpy as np
import pytest

# camera parameters of the reference's configs: config/euroc_vio.yaml:11-18, config/tum_vi.yaml:11-19
EUROC_L = ([458.654, 457.296, 367.215, 248.375], [-0.28340811, 0.07395907, 0.00019359, 1.76187114e-05])
EUROC_R = ([457.587, 456.134, 379.999, 255.238], [-0.28368365, 0.07451284, -0.00010473, -3.55590700e-05])
TUMVI_L = ([191.75556798912652, 191.74816751185256, 254.9226487139376, 256.8780365577954],
           [0.6246288732884442, 1.0598071085569876])


def grid(w, h, step):
    y, x = np.mgrid[0:h:step, 0:w:step]
    return np.stack([x.ravel(), y.ravel()], 1).astype(np.float64) + 0.25


def test_oracle_round_trip_and_known_points(oracle):
    for model, (K, D), (w, h) in ((1, EUROC_L, (752, 480)), (1, EUROC_R, (752, 480)), (2, TUMVI_L, (512, 512))):
        P = K + D + [0.0] * (9 - len(K) - len(D))
        uv = grid(w, h, 23)
        xy = oracle.unproject(model, P, uv)
        ok = np.isfinite(xy).all(axis=1)
        assert ok.mean() > (0.99 if model == 1 else 0.6)          # the fisheye's corners look backwards
        assert np.abs(oracle.project(model, P, xy[ok]) - uv[ok]).max() < 1e-9
        centre = oracle.unproject(model, P, [[K[2], K[3]]])
        assert np.abs(centre).max() < 1e-15                        # the principal point is the optical axis
    # no distortion: plain pinhole
    xy = oracle.unproject(1, [400.0, 300.0, 320.0, 240.0, 0, 0, 0, 0, 0], [[720.0, 540.0]])
    assert np.allclose(xy, [[1.0, 1.0]], atol=1e-15)
    # EUCM with alpha = 0 is a pinhole as well
    xy = oracle.unproject(2, [400.0, 300.0, 320.0, 240.0, 0.0, 1.0, 0, 0, 0], [[720.0, 540.0]])
    assert np.allclose(xy, [[1.0, 1.0]], atol=1e-15)


@pytest.mark.gpu
def test_gpu_unproject_matches_oracle(rsvio, oracle):
    for model, (K, D), (w, h) in ((1, EUROC_L, (752, 480)), (1, EUROC_R, (752, 480)), (2, TUMVI_L, (512, 512))):
        cm = rsvio.CameraModel.pinhole_radtan(K, D) if model == 1 else rsvio.CameraModel.eucm(K, D)
        uv = grid(w, h, 7).astype(np.float32)
        got = rsvio.unproject_points(cm, uv)
        want = oracle.unproject(model, list(cm.params), uv.astype(np.float64))
        assert np.array_equal(np.isfinite(got).all(axis=1), np.isfinite(want).all(axis=1))
        ok = np.isfinite(want).all(axis=1)
        # both sides solve the same equations in f64 (different iterations); the output is rounded to f32
        err = np.abs(got[ok].astype(np.float64) - want[ok]).max()
        assert err <= 2e-7 * (1.0 + np.abs(want[ok]).max())


@pytest.mark.gpu
def test_tracker_exports_undistorted_coordinates(rsvio, oracle):
    from rsvio_b200.synth import StereoSequence
    w, h = 376, 240
    seq = StereoSequence("C1", w=w, h=h)
    K = [229.3, 228.6, 183.6, 124.2]
    D = [-0.2834, 0.0739, 0.00019, 1.76e-05]
    tr = rsvio.StereoPatchTracker(40, 20, 0.01, width=w, height=h, levels=3)
    tr.process_frame(*seq.frame(0))
    assert tr.undistorted_view() is None                          # no model: the reference's (-1, -1) placeholder

    class Frame:
        def __init__(self):
            self.left, self.right = [], []

        def add_left_feature(self, f):
            self.left.append(f)

        def add_right_feature(self, f):
            self.right.append(f)

    f0 = Frame()
    tr.process_frame(*seq.frame(1), f0)
    assert f0.left and all(f.undistorted_coord == (-1.0, -1.0) for f in f0.left)
    tr.set_camera(0, rsvio.CameraModel.pinhole_radtan(K, D))
    tr.set_camera(1, rsvio.CameraModel.pinhole_radtan([k + 1.0 for k in K], D))
    f1 = Frame()
    tr.process_frame(*seq.frame(2), f1)
    assert len(f1.left) > 20 and len(f1.left) == len(f1.right)
    for feats, Kc in ((f1.left, K), (f1.right, [k + 1.0 for k in K])):
        uv = np.array([f.pixel_coord for f in feats], np.float64)
        want = oracle.unproject(1, Kc + D + [0.0], uv)
        got = np.array([f.undistorted_coord for f in feats], np.float64)
        assert np.abs(got - want).max() <= 2e-7 * (1.0 + np.abs(want).max())
    # remove_id keeps records and coordinates aligned
    victim = f1.left[3].feature_id
    tr.remove_id([victim])
    rec, cnt = tr.result_view()
    uxy = tr.undistorted_view()
    assert victim not in rec["id"][0, :cnt[0]]
    uv = np.stack([rec["x"][0, :cnt[0]], rec["y"][0, :cnt[0]]], 1).astype(np.float64)
    assert np.abs(uxy[0, :cnt[0]] - oracle.unproject(1, K + D + [0.0], uv)).max() <= 2e-7 * 3
    tr.set_camera(0, rsvio.CameraModel(rsvio.CAMERA_NONE))
    tr.set_camera(1, rsvio.CameraModel(rsvio.CAMERA_NONE))
    tr.process_frame(*seq.frame(3))
    assert tr.undistorted_view() is None


# ---- ingest row f-3: decoded frames (DynamicImage::to_luma8 of the dataset players) -------------------------
def test_oracle_to_luma8_known_values(oracle):
    g16 = np.array([[0, 128, 129, 257, 65535, 32768, 385, 386]], np.uint16)
    assert oracle.to_luma8(g16, oracle.PIXEL_LUMA16).tolist() == [[0, 0, 1, 1, 255, 128, 1, 2]]   # (c + 128) // 257
    rgb = np.array([[[255, 255, 255], [255, 0, 0], [0, 255, 0], [0, 0, 255], [10, 20, 30]]], np.uint8)
    want = [[255, 2126 * 255 // 10000, 7152 * 255 // 10000, 722 * 255 // 10000,
             (2126 * 10 + 7152 * 20 + 722 * 30) // 10000]]
    assert oracle.to_luma8(rgb, oracle.PIXEL_RGB8).tolist() == want


@pytest.mark.gpu
def test_gpu_to_luma8_and_decoded_frames(rsvio, oracle):
    from rsvio_b200.synth import StereoSequence
    rng = np.random.default_rng(5)
    g16 = rng.integers(0, 65536, (97, 131), dtype=np.uint16)
    rgb = rng.integers(0, 256, (97, 131, 3), dtype=np.uint8)
    assert np.array_equal(rsvio.convert_to_luma8(g16, rsvio.PIXEL_LUMA16), oracle.to_luma8(g16, oracle.PIXEL_LUMA16))
    assert np.array_equal(rsvio.convert_to_luma8(rgb, rsvio.PIXEL_RGB8), oracle.to_luma8(rgb, oracle.PIXEL_RGB8))
    # a tracker fed 16-bit / RGB frames == a tracker fed their to_luma8()
    w, h = 376, 240
    seq = StereoSequence("C1", w=w, h=h)
    for fmt, ofmt in ((rsvio.PIXEL_LUMA16, oracle.PIXEL_LUMA16), (rsvio.PIXEL_RGB8, oracle.PIXEL_RGB8)):
        a = rsvio.StereoPatchTracker(40, 20, 0.01, width=w, height=h, levels=3)
        b = rsvio.StereoPatchTracker(40, 20, 0.01, width=w, height=h, levels=3)
        for t in range(3):
            l8, r8 = seq.frame(t)
            if fmt == rsvio.PIXEL_LUMA16:
                lraw = l8.astype(np.uint16) * 257 + rng.integers(-100, 100, l8.shape).clip(0).astype(np.uint16)
                rraw = r8.astype(np.uint16) * 257
            else:
                lraw = np.stack([l8, l8, l8], 2)
                rraw = np.stack([r8, (r8 // 2 + 60).astype(np.uint8), r8], 2)
            a.process_decoded(lraw, rraw, fmt)
            b.process_frame(oracle.to_luma8(lraw, ofmt), oracle.to_luma8(rraw, ofmt))
            for cam in (0, 1):
                ia, Ta = a.transforms(cam)
                ib, Tb = b.transforms(cam)
                assert np.array_equal(ia, ib) and np.array_equal(Ta, Tb)
        assert a.num_tracks(0) > 20


@pytest.mark.gpu
def test_gpu_matches_golden_rows_f(rsvio):
    import os
    from util import GOLDEN_DIR
    g = np.load(os.path.join(GOLDEN_DIR, "golden_rows_f.npz"))
    for name, mk in (("euroc", rsvio.CAMERA_PINHOLE_RADTAN), ("tumvi", rsvio.CAMERA_EUCM)):
        P = g[f"unproject_{name}_params"]
        cm = rsvio.CameraModel(mk, 0, (rsvio.binding.C.c_double * 9)(*P))
        got = rsvio.unproject_points(cm, g[f"unproject_{name}_uv"]).astype(np.float64)
        want = g[f"unproject_{name}_xy"]
        ok = np.isfinite(want).all(axis=1)
        assert np.array_equal(np.isfinite(got).all(axis=1), ok)
        assert np.abs(got[ok] - want[ok]).max() <= 2e-7 * (1.0 + np.abs(want[ok]).max())
    assert np.array_equal(rsvio.convert_to_luma8(g["luma16_in"], rsvio.PIXEL_LUMA16), g["luma16_out"])
    assert np.array_equal(rsvio.convert_to_luma8(g["rgb8_in"], rsvio.PIXEL_RGB8), g["rgb8_out"])
