"""Per-stage device times of one stereo stream (C1-like, 752x480) -- latency view of the path.
Usage: python tools/single_stream_stages.py [levels] [grid]"""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "rs-vio_b200"))
import numpy as np
import rsvio_b200
from rsvio_b200 import synth

levels = int(sys.argv[1]) if len(sys.argv) > 1 else 3
grid = int(sys.argv[2]) if len(sys.argv) > 2 else 40
w, h = 752, 480
ss = synth.StereoSequence("C1")
seq = [ss.frame(t) for t in range(12)]
tr = rsvio_b200.StereoPatchTracker(grid, 20, 0.01, width=w, height=h, levels=levels)
tr.set_profiling(True)
import time
acc = {}
n = 0
for rep in range(20):
    for k, (l, r) in enumerate(seq):
        t0 = time.perf_counter()
        tr.process_frame(l, r)
        dt = time.perf_counter() - t0
        if rep >= 2:
            st = tr.stage_times()
            for kk, v in st.items():
                acc[kk] = acc.get(kk, 0.0) + v[0] / max(1, 1)
            acc["wall"] = acc.get("wall", 0.0) + dt * 1e3
            n += 1
print({k: round(v / n, 4) for k, v in acc.items()}, "tracks", tr.num_tracks())
# stage_times returns accumulated ms since the last reset: one frame per call here
