"""Pyramid kernel microbenchmark (device-resident batches): prints GB/s of algorithmic bytes.
Used under ncu for the per-kernel launch list:  python tools/pyr_bench.py [C1|C4|C2|C3] [n_images] [repeats]"""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "rs-vio_b200"))
import torch  # noqa: E402
import rsvio_b200  # noqa: E402

SHAPES = {"C1": (752, 480, 3), "C2": (512, 512, 4), "C3": (800, 400, 6), "C4": (1920, 1200, 5), "EUROC6": (752, 480, 6)}


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "C1"
    w, h, lv = SHAPES[name]
    n = int(sys.argv[2]) if len(sys.argv) > 2 else (128 if name != "C4" else 64)
    rep = int(sys.argv[3]) if len(sys.argv) > 3 else 20
    L = rsvio_b200.load_library()
    src = torch.randint(0, 256, (n, h, w), dtype=torch.uint8, device="cuda")
    pb = L.rsvio_pyramid_device_bytes(w, h, lv)
    dst = torch.empty(n * pb, dtype=torch.uint8, device="cuda")
    ms = C.c_float(0)
    for r in (3, rep):
        st = L.rsvio_build_pyramid_u8_device(0, C.c_void_p(src.data_ptr()), n, w, h, w, w * h, lv,
                                             C.c_void_p(dst.data_ptr()), r, C.byref(ms))
        assert st == 0, st
    ab = n * (w * h + rsvio_b200.pyramid_bytes(w, h, lv))
    # cold inputs: rotate over enough distinct batches that none is still in the 126 MB L2, one launch per call
    nb = max(2, int(400e6 // ab) + 1)
    srcs = [torch.randint(0, 256, (n, h, w), dtype=torch.uint8, device="cuda") for _ in range(nb)]
    dsts = [torch.empty(n * pb, dtype=torch.uint8, device="cuda") for _ in range(nb)]
    cold = []
    for i in range(3 * nb):
        st = L.rsvio_build_pyramid_u8_device(0, C.c_void_p(srcs[i % nb].data_ptr()), n, w, h, w, w * h, lv,
                                             C.c_void_p(dsts[i % nb].data_ptr()), 1, C.byref(ms))
        assert st == 0, st
        if i >= nb:
            cold.append(ms.value)
    cold_ms = sum(cold) / len(cold)
    print(f"{name} x{n} cold ({nb} rotating batches, 1 launch per call): {cold_ms * 1e3:.1f} us, "
          f"{ab / (cold_ms * 1e-3) / 1e9:.0f} GB/s")
    print(f"{name} x{n}: {ms.value * 1e3:.1f} us/launch-set, {ab / 1e6:.1f} MB algorithmic, "
          f"{ab / (ms.value * 1e-3) / 1e9:.0f} GB/s")


if __name__ == "__main__":
    main()
