"""Row f-2: per-frame time of the crate B tracker (rsvio_b_tracker_process_frame) on 752x480 frames of the C1 sequence
with the reference's config.yaml settings.  Under ncu (--metrics gpu__time_duration.sum) it gives the per-kernel list.
Usage: python tools/crate_b_probe.py [frames]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "rs-vio_b200"))
import numpy as np  # noqa: E402
import rsvio_b200  # noqa: E402
from rsvio_b200.synth import StereoSequence  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 30
    seq = StereoSequence("C1", stream_id=0)
    imgs = [(seq.frame(t)[0].astype(np.float32) / np.float32(255.0)) for t in range(8)]
    tr = rsvio_b200.FeatureTrackerB(752, 480, detection_threshold=0.02)
    dev, wall, feats = [], [], []
    for i in range(n):
        k = i % 14
        img = imgs[k if k < 8 else 14 - k]
        t0 = time.perf_counter()
        f = tr.process_frame(img, i)
        wall.append(time.perf_counter() - t0)
        dev.append(tr.last_frame_ms())
        feats.append((len(f), tr.n_new))
    print("features (total, new) per frame:", feats[:6], "...")
    print(f"device ms per frame: median {np.median(dev[3:]):.3f}, wall ms: median {1e3 * np.median(wall[3:]):.3f}")


if __name__ == "__main__":
    main()
