"""Per-stage device times of the 64-stream C5 step (profiling mode: one stream, CUDA events between the stages),
once with device-resident frames and once with pinned host frames (upload beside the kernels).
Usage: python tools/stage_probe.py [streams] [steps]"""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "rs-vio_b200")]
import numpy as np  # noqa: E402
import torch  # noqa: E402
import rsvio_b200  # noqa: E402
from bench import generate_frames, ring_index, RING  # noqa: E402

S = int(sys.argv[1]) if len(sys.argv) > 1 else 64
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 40
w, h = 752, 480
cache = f"/tmp/rsvio_frames_C5_{S}.npy"   # A/B runs on one box: generate the synthetic frames once
if os.path.exists(cache):
    frames = np.load(cache)
else:
    frames = generate_frames("C5", list(range(S)), RING, 16)
    np.save(cache, frames)
host = torch.from_numpy(frames).pin_memory()
dev = host.cuda()
tr = rsvio_b200.StereoPatchTracker(40, 20, 0.01, width=w, height=h, levels=3, n_streams=S)
ptrs = [((C.c_void_p * S)(*[host[k, 0, s].data_ptr() for s in range(S)]),
         (C.c_void_p * S)(*[host[k, 1, s].data_ptr() for s in range(S)])) for k in range(RING)]


def run(kind, first, n):
    def submit(i):
        k = ring_index(i)
        if kind == "device":
            tr.submit_device(dev[k, 0].data_ptr(), dev[k, 1].data_ptr(), w, w * h)
        else:
            tr.submit_batch_ptr(ptrs[k][0], ptrs[k][1], w)
    submit(first)
    for i in range(first + 1, first + n):
        submit(i)
        tr.sync()
    tr.sync()


for kind in (("device",) if os.environ.get("RSVIO_PROBE_DEVICE_ONLY") else ("device", "host")):
    tr.reset()
    run(kind, 0, 8)
    tr.set_profiling(True)
    tr.stage_times(reset=True)
    run(kind, 8, steps)
    st = tr.stage_times(reset=True)
    tr.set_profiling(False)
    print(kind, {k: round(1e3 * v[0] / steps, 1) for k, v in st.items()}, "us per step")
