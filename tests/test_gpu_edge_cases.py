"""GPU: edge cases through the C ABI -- empty and ragged inputs, strides and pitches, extreme sizes,
pinned vs pageable hand-off, stale ids."""
import ctypes as C

import numpy as np
import pytest

from util import identity_T6, random_image, scene_pair

pytestmark = pytest.mark.gpu


def test_track_points_empty_and_single(rsvio, oracle):
    a, b = scene_pair(160, 120, 3)
    pa, pb = oracle.build_pyramid(a, 3), oracle.build_pyramid(b, 3)
    T, st, it = rsvio.track_points(pa, pb, 160, 120, 3, np.zeros((0, 6), np.float32), 20, 0.01)
    assert T.shape == (0, 6) and st.shape == (0,)
    T, st, it = rsvio.track_points(pa, pb, 160, 120, 3, identity_T6([[80.0, 60.0]]), 20, 0.01)
    Tr, sr, _, _ = oracle.track_points(pa, pb, 160, 120, 3, identity_T6([[80.0, 60.0]]), 20, 0.01)
    assert st[0] == sr[0] and (st[0] != 0 or np.abs(T - Tr).max() < 1e-3)


def test_max_iterations_zero_and_nan_inputs(rsvio, oracle):
    a, b = scene_pair(160, 120, 4)
    pa, pb = oracle.build_pyramid(a, 3), oracle.build_pyramid(b, 3)
    T0 = identity_T6([[80.0, 60.0], [np.nan, 10.0], [40.0, np.inf], [-1e30, 5.0]])
    for max_it in (0, 1):
        T, st, it = rsvio.track_points(pa, pb, 160, 120, 3, T0, max_it, 0.01)
        Tr, sr, ir, _ = oracle.track_points(pa, pb, 160, 120, 3, T0, max_it, 0.01)
        assert np.array_equal(st == 0, sr == 0)
        assert np.array_equal(it, ir)
    assert (st[1:] != 0).all()


def test_pyramid_with_stride_and_pitch(rsvio, oracle):
    """rows padded to a larger stride and images a larger pitch apart (unaligned: generic path; aligned: TMA path)"""
    w, h, levels, n = 752, 480, 3, 3
    L = rsvio.load_library()
    for stride, pad_rows in ((752 + 16, 2), (752 + 5, 1)):
        pitch = stride * (h + pad_rows)
        buf = np.zeros(n * pitch, np.uint8)
        imgs = []
        for j in range(n):
            im = random_image(w, h, 40 + j)
            imgs.append(im)
            view = buf[j * pitch:j * pitch + stride * h].reshape(h, stride)
            view[:, :w] = im
            view[:, w:] = 255 - (j + 1)   # padding must never leak into the result
        out = np.empty((n, rsvio.pyramid_bytes(w, h, levels)), np.uint8)
        st = L.rsvio_build_pyramid_u8(0, buf, n, w, h, stride, pitch, levels, out)
        assert st == 0
        for j in range(n):
            assert np.array_equal(out[j], oracle.build_pyramid(imgs[j], levels))


@pytest.mark.parametrize("w,h,levels", [(16, 16, 1), (16, 16, 3), (24, 8, 2), (65535, 8, 2), (8, 4096, 4)])
def test_pyramid_extreme_shapes(rsvio, oracle, w, h, levels):
    img = random_image(w, h, 7)
    assert np.array_equal(rsvio.build_pyramid(img, levels)[0], oracle.build_pyramid(img, levels))


def test_invalid_arguments_return_status(rsvio):
    L = rsvio.load_library()
    img = np.zeros((64, 64), np.uint8)
    out = np.zeros(64 * 64 * 2, np.uint8)
    assert L.rsvio_build_pyramid_u8(0, img, 1, 64, 64, 32, 0, 2, out) == -1        # stride < width
    assert L.rsvio_build_pyramid_u8(0, img, 1, 64, 64, 64, 0, 9, out) == -1        # too many levels
    assert L.rsvio_build_pyramid_u8(99, img, 1, 64, 64, 64, 0, 2, out) == -2       # no such device
    with pytest.raises(rsvio.RsvioError):
        rsvio.StereoPatchTracker(40, 20, 0.01, width=752, height=480, levels=12)
    with pytest.raises(rsvio.RsvioError):
        rsvio.StereoPatchTracker(3, 20, 0.01, width=752, height=480, levels=3)     # grid smaller than a FAST circle


def test_featureless_images_yield_no_tracks(rsvio):
    tr = rsvio.StereoPatchTracker(40, 20, 0.01, width=376, height=240, levels=3)
    flat = np.full((240, 376), 100, np.uint8)
    for _ in range(3):
        tr.process_frame(flat, flat)
        assert tr.num_tracks(0) == 0 and tr.num_tracks(1) == 0
    assert tr.get_track_points() == [{}, {}]
    tr.remove_id(np.array([1, 2, 3], np.uint64))       # nothing to remove: no error
    assert tr.counters()["frames"] == 3


def test_remove_unknown_ids_is_a_noop(rsvio):
    from rsvio_b200.synth import StereoSequence
    seq = StereoSequence("C1", w=376, h=240)
    tr = rsvio.StereoPatchTracker(40, 20, 0.01, width=376, height=240, levels=3)
    tr.process_frame(*seq.frame(0))
    before = tr.tracks(0)[0]
    tr.remove_id(np.array([10**9, 10**12], np.uint64))
    assert np.array_equal(tr.tracks(0)[0], before)
    tr.remove_id(before)                                 # remove everything
    assert tr.num_tracks(0) == 0 and tr.num_tracks(1) == 0
    tr.process_frame(*seq.frame(1))                      # all cells empty again: new ids continue after the old ones
    ids = tr.tracks(0)[0]
    assert len(ids) > 0 and ids.min() > before.max()


def test_pinned_and_pageable_inputs_agree(rsvio):
    from rsvio_b200.synth import StereoSequence
    w, h = 376, 240
    seq = StereoSequence("C1", w=w, h=h)
    frames = [seq.frame(t) for t in range(3)]
    L = rsvio.load_library()
    ptr = C.c_void_p()
    assert L.rsvio_host_alloc(C.byref(ptr), 2 * w * h) == 0
    pinned = np.frombuffer((C.c_uint8 * (2 * w * h)).from_address(ptr.value), np.uint8).reshape(2, h, w)
    a = rsvio.StereoPatchTracker(40, 20, 0.01, width=w, height=h, levels=3)
    b = rsvio.StereoPatchTracker(40, 20, 0.01, width=w, height=h, levels=3)
    for l, r in frames:
        a.process_frame(l, r)                             # pageable numpy memory: staged copy
        pinned[0], pinned[1] = l, r
        b.process_frame(pinned[0], pinned[1])             # pinned: DMA straight from the caller's buffer
        for cam in (0, 1):
            ia, Ta = a.transforms(cam)
            ib, Tb = b.transforms(cam)
            assert np.array_equal(ia, ib) and np.array_equal(Ta, Tb)
    del a, b, pinned
    assert L.rsvio_host_free(ptr) == 0


def test_many_small_streams_in_one_batch(rsvio):
    """ragged batch: streams with very different numbers of live tracks share the same launches"""
    from rsvio_b200.synth import StereoSequence
    S, w, h = 5, 376, 240
    seqs = [StereoSequence("C5", stream_id=s, w=w, h=h) for s in range(S)]
    flat = np.full((h, w), 90, np.uint8)
    batch = rsvio.StereoPatchTracker(40, 20, 0.01, width=w, height=h, levels=3, n_streams=S)
    single = rsvio.StereoPatchTracker(40, 20, 0.01, width=w, height=h, levels=3)
    for t in range(3):
        fr = [q.frame(t) for q in seqs]
        fr[1] = (flat, flat)                               # stream 1 never has a track
        fr[3] = (fr[3][0], flat)                           # stream 3: stereo always fails
        batch.process_batch([f[0] for f in fr], [f[1] for f in fr])
        single.process_frame(*fr[4])
    assert batch.num_tracks(0, 1) == 0 and batch.num_tracks(0, 3) == 0
    assert batch.num_tracks(0, 0) > 20
    ib, Tb = batch.transforms(0, 4)
    i1, T1 = single.transforms(0)
    assert np.array_equal(ib, i1) and np.array_equal(Tb, T1)


def test_calls_leave_the_current_device_alone(rsvio):
    """a handle on device 0 driven from a thread whose current device is another one (needs >= 2 GPUs;
    on a 1-GPU box this only checks that the guard is a no-op)"""
    import torch
    from rsvio_b200.synth import StereoSequence
    n = torch.cuda.device_count()
    other = 1 if n > 1 else 0
    seq = StereoSequence("C1", w=256, h=160)
    tr = rsvio.StereoPatchTracker(40, 20, 0.01, width=256, height=160, levels=3, device=0)
    torch.cuda.set_device(other)
    for t in range(2):
        tr.process_frame(*seq.frame(t))
        assert torch.cuda.current_device() == other
    assert tr.num_tracks(0) > 0
    pyr = rsvio.build_pyramid(seq.frame(0)[0], 3, device=0)
    assert pyr.shape[0] == 1 and torch.cuda.current_device() == other
    torch.cuda.set_device(0)


def test_two_handles_driven_from_two_threads(rsvio):
    """distinct handles are independent (own stream, own tables, no global mutable state): two threads that
    drive one handle each get the results of a single-threaded run (ctypes releases the GIL during the calls)"""
    import threading
    from rsvio_b200.synth import StereoSequence
    w, h, frames = 256, 160, 5
    seqs = [StereoSequence("C5", stream_id=s, w=w, h=h) for s in range(2)]
    data = [[q.frame(t) for t in range(frames)] for q in seqs]
    want = []
    for s in range(2):
        tr = rsvio.StereoPatchTracker(40, 20, 0.01, width=w, height=h, levels=3)
        for l, r in data[s]:
            tr.process_frame(l, r)
        want.append([tr.transforms(c) for c in (0, 1)])
    got, errors = [None, None], []

    def work(s):
        try:
            tr = rsvio.StereoPatchTracker(40, 20, 0.01, width=w, height=h, levels=3)
            for _ in range(3):                       # several passes: more chances to interleave
                tr.reset()
                for l, r in data[s]:
                    tr.process_frame(l, r)
            got[s] = [tr.transforms(c) for c in (0, 1)]
        except Exception as e:                       # pragma: no cover
            errors.append(e)

    threads = [threading.Thread(target=work, args=(s,)) for s in range(2)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors
    for s in range(2):
        for cam in (0, 1):
            assert np.array_equal(got[s][cam][0], want[s][cam][0])
            assert np.array_equal(got[s][cam][1], want[s][cam][1])


def test_capacity_overflow_is_reported_per_frame_not_forever(rsvio):
    """a table that is too small: the frame that overflows returns RSVIO_ERR_CAPACITY through the C ABI (never
    silently); it is NOT fatal -- the frame is processed, the points that fitted are kept, the mirrors deliver the
    frame and warn (the reference's maps are unbounded, its process_frame cannot fail); a later frame that does not
    overflow returns ok, the overflow count stays queryable"""
    from rsvio_b200.synth import StereoSequence
    w, h = 376, 240
    seq = StereoSequence("C1", w=w, h=h)
    tr = rsvio.StereoPatchTracker(40, 20, 0.01, width=w, height=h, levels=3, max_tracks=20)
    for t in range(2):
        with pytest.warns(rsvio.CapacityWarning):
            tr.process_frame(*seq.frame(t))      # ~50 cells want a point, room for 20
        n0 = tr.num_tracks(0)
        assert 0 < n0 <= 20 and tr.num_tracks(1) == n0
        assert tr.overflow_frames() == t + 1
    flat = np.full((h, w), 77, np.uint8)
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("error")
        tr.process_frame(flat, flat)             # every track dies, nothing is detected: no overflow, no warning
    assert tr.num_tracks(0) == 0 and tr.overflow_frames() == 2
    tr.process_frame(flat, flat)
