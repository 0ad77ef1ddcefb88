"""Per-stage device times of ONE stereo stream of a BASELINE config (C1..C4, EUROC_SHIPPED): device-resident frames,
profiling mode (one stream, CUDA events between the stages).
Usage: python tools/config_stages.py <config> [steps]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "rs-vio_b200")]
import torch  # noqa: E402
import rsvio_b200  # noqa: E402
from bench import CONFIGS, RING, generate_frames, ring_index  # noqa: E402

cfg = sys.argv[1] if len(sys.argv) > 1 else "C4"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 24
p = CONFIGS[cfg]
w, h = p["w"], p["h"]
dev = torch.from_numpy(generate_frames(cfg, [0], RING, 1)).cuda()
tr = rsvio_b200.StereoPatchTracker(p["grid"], p["max_it"], p["thr"], width=w, height=h, levels=p["levels"], n_streams=1)


def run(first, n):
    for i in range(first, first + n):
        k = ring_index(i)
        tr.submit_device(dev[k, 0].data_ptr(), dev[k, 1].data_ptr(), w, w * h)
        tr.sync()


run(0, 8)
tr.set_profiling(True)
tr.stage_times(reset=True)
run(8, steps)
st = tr.stage_times(reset=True)
tr.set_profiling(False)
us = {k: round(1e3 * v[0] / steps, 1) for k, v in st.items()}
print(cfg, "tracks", tr.num_tracks(0) + tr.num_tracks(1), "sum", round(sum(us.values()), 1), us, "us per frame")
