"""Cost of the fused unprojection (row f-1): 'download' stage (record export kernel + flag copy) of a 64-stream
handle with and without camera models, and the stateless kernel on 1 M points.
Usage: python tools/unproject_probe.py"""
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

EUROC_L = ([458.654, 457.296, 367.215, 248.375], [-0.28340811, 0.07395907, 0.00019359, 1.76187114e-05])
EUROC_R = ([457.587, 456.134, 379.999, 255.238], [-0.28368365, 0.07451284, -0.00010473, -3.55590700e-05])


def main():
    p = CONFIGS["C5"]
    w, h, S, ring = p["w"], p["h"], 64, 8
    frames = bench.generate_frames("C5", list(range(S)), ring, 16)
    dev = torch.device("cuda", 0)
    devbuf = torch.from_numpy(frames).to(dev)
    torch.cuda.synchronize()
    for with_cam in (False, True):
        tr = rsvio_b200.StereoPatchTracker(p["grid"], p["max_it"], p["thr"], width=w, height=h, levels=p["levels"], n_streams=S)
        if with_cam:
            tr.set_camera(0, rsvio_b200.CameraModel.pinhole_radtan(*EUROC_L))
            tr.set_camera(1, rsvio_b200.CameraModel.pinhole_radtan(*EUROC_R))

        def run(first, n):
            for i in range(first, first + n):
                k = bench.ring_index(i)
                tr.submit_device(devbuf[k, 0].data_ptr(), devbuf[k, 1].data_ptr(), w, w * h)
                tr.sync()
        run(0, 10)
        tr.set_profiling(True)
        tr.stage_times(reset=True)
        t0 = time.perf_counter()
        run(10, 50)
        dt = (time.perf_counter() - t0) / 50
        st = tr.stage_times(reset=True)
        n = int(tr.result_view()[1].sum())
        print(f"camera models {'on ' if with_cam else 'off'}: export stage {st['download'][0] / 50 * 1e3:.1f} us per step "
              f"({n} records), {dt * 1e3:.3f} ms per synchronous step")
    cm = rsvio_b200.CameraModel.pinhole_radtan(*EUROC_L)
    xy = np.random.default_rng(0).uniform([0, 0], [752, 480], (1 << 20, 2)).astype(np.float32)
    rsvio_b200.unproject_points(cm, xy)
    t0 = time.perf_counter()
    out = rsvio_b200.unproject_points(cm, xy)
    print(f"stateless: 1 M points in {(time.perf_counter() - t0) * 1e3:.2f} ms wall incl. copies, finite {np.isfinite(out).all()}")


if __name__ == "__main__":
    main()
