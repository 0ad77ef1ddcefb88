"""Probe: 64 streams as H handles of 64/H streams each, driven round-robin from one thread (device-resident
frames).  Shows how much of the latency-bound tail of a frame can hide behind another handle's LK launch.
Usage: python tools/two_handles_probe.py [steps] [H:B ...]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "rs-vio_b200"))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
import rsvio_b200  # noqa: E402
from rsvio_b200.synth import CONFIGS  # noqa: E402
import bench  # noqa: E402


def main():
    """argv: steps, then any number of H:B pairs (handles : persistent LK CTAs per SM, RSVIO_LK_BLOCKS_PER_SM)"""
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 100
    combos = [tuple(int(x) for x in a.split(":")) for a in sys.argv[2:]] or [(1, 8), (2, 8), (2, 7), (2, 6), (1, 7)]
    S_total, ring = 64, 8
    p = CONFIGS["C5"]
    w, h = p["w"], p["h"]
    frames = bench.generate_frames("C5", list(range(S_total)), ring, 16)     # [ring][cam][S][h][w]
    dev = torch.device("cuda", 0)
    devbuf = torch.from_numpy(frames).to(dev)
    torch.cuda.synchronize()
    for H, blocks in combos:
        os.environ["RSVIO_LK_BLOCKS_PER_SM"] = str(blocks)   # read by the library at every LK launch
        S = S_total // H
        trs = [rsvio_b200.StereoPatchTracker(p["grid"], p["max_it"], p["thr"], width=w, height=h, levels=p["levels"],
                                             n_streams=S) for _ in range(H)]

        def submit(i):
            k = i % (2 * ring - 2)
            k = k if k < ring else 2 * ring - 2 - k
            for j, tr in enumerate(trs):
                tr.submit_device(devbuf[k, 0, j * S].data_ptr(), devbuf[k, 1, j * S].data_ptr(), w, w * h)

        def run(first, n):
            submit(first)
            for i in range(first + 1, first + n):
                submit(i)
                for tr in trs:
                    tr.sync()
            for tr in trs:
                tr.sync()

        run(0, 10)
        torch.cuda.synchronize()
        c0 = sum(tr.counters()["tracks"] for tr in trs)
        t0 = time.perf_counter()
        run(10, steps)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        c1 = sum(tr.counters()["tracks"] for tr in trs)
        print(f"{H} handle(s) x {S} streams, {blocks} LK CTAs/SM: {dt / steps * 1e3:.4f} ms per 64-stream step, "
              f"{(c1 - c0) / dt / 1e6:.1f} M tracks/s", flush=True)
        for tr in trs:
            tr.close()


if __name__ == "__main__":
    main()
