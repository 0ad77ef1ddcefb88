"""Soak: a 16-stream handle over many frames with uploads from pinned memory; checks that track counts stay in
range, ids keep increasing, no status errors, and device memory does not grow.
Usage: python tools/soak.py [frames] [streams] [device]   ("device": frames resident in HBM, submit_device -- with >= 24
streams that is the split temporal LK pipeline running flat out)"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "rs-vio_b200"))
sys.path.insert(0, ROOT)
import ctypes as C  # noqa: E402
import numpy as np  # noqa: E402
import torch  # noqa: E402
import rsvio_b200  # noqa: E402
from rsvio_b200.synth import CONFIGS  # noqa: E402
import bench  # noqa: E402


def main():
    n_frames = int(sys.argv[1]) if len(sys.argv) > 1 else 3000
    p = CONFIGS["C5"]
    w, h, S, ring = p["w"], p["h"], (int(sys.argv[2]) if len(sys.argv) > 2 else 16), 8
    on_device = len(sys.argv) > 3 and sys.argv[3] == "device"
    frames = bench.generate_frames("C5", list(range(S)), ring, 16)
    host = torch.from_numpy(frames).pin_memory()
    torch.cuda.init()
    tr = rsvio_b200.StereoPatchTracker(p["grid"], p["max_it"], p["thr"], width=w, height=h, levels=p["levels"], n_streams=S)
    tr.set_camera(0, rsvio_b200.CameraModel.pinhole_radtan([458.654, 457.296, 367.215, 248.375], [-0.2834, 0.0739, 0.0002, 1.8e-05]))
    ptrs = [((C.c_void_p * S)(*[host[k, 0, s].data_ptr() for s in range(S)]),
             (C.c_void_p * S)(*[host[k, 1, s].data_ptr() for s in range(S)])) for k in range(ring)]
    free0 = None
    last_max_id = -1
    dev = host.cuda() if on_device else None

    def submit(k):
        if on_device:
            tr.submit_device(dev[k, 0].data_ptr(), dev[k, 1].data_ptr(), w, w * h)
        else:
            tr.submit_batch_ptr(*ptrs[k])

    submit(0)
    for i in range(1, n_frames):
        submit(bench.ring_index(i))
        tr.sync()
        if i % 500 == 0 or i == 50:
            rec, cnt = tr.result_view()
            uxy = tr.undistorted_view()
            free, _ = torch.cuda.mem_get_info()
            free0 = free if free0 is None else free0
            mx = max(int(rec["id"][g, :cnt[g]].max()) for g in range(S) if cnt[g])
            assert mx >= last_max_id
            last_max_id = mx
            assert 50 * S < cnt[:S].sum() < 400 * S, cnt[:S]
            assert np.isfinite(uxy[0, :cnt[0]]).all()
            assert abs(free - free0) < (64 << 20), (free, free0)
            print(f"frame {i}: {int(cnt.sum())} tracks, max id {mx}, free {free >> 20} MiB", flush=True)
    tr.sync()
    print("soak ok", tr.counters())


if __name__ == "__main__":
    main()
