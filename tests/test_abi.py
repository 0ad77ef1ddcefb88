"""CPU: the C-ABI library loads and exports every symbol include/rsvio_b200.h declares; without a
GPU it fails loudly instead of falling back.  No compute calls here."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_functions():
    src = open(os.path.join(ROOT, "include", "rsvio_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(rsvio_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_are_exported(rsvio):
    lib = ctypes.CDLL(rsvio.library_path())
    declared = _declared_functions()
    assert len(declared) >= 25
    missing = [s for s in declared if not hasattr(lib, s)]
    assert not missing, missing
    assert set(declared) == set(rsvio.ABI_SYMBOLS)
    assert lib.rsvio_abi_version() == 1


def test_product_does_not_link_or_import_the_oracle(rsvio):
    # the product path must never route through oracle/
    import subprocess
    out = subprocess.run(["ldd", rsvio.library_path()], capture_output=True, text=True).stdout
    assert "oracle" not in out
    pkg = os.path.join(ROOT, "rs-vio_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".hpp", ".rs")):
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "import oracle" not in txt and "from oracle" not in txt and "rsvio_oracle" not in txt, f


def test_invalid_arguments_are_rejected_without_touching_the_gpu(rsvio):
    L = rsvio.load_library()
    from rsvio_b200.binding import Config
    cfg = Config()
    L.rsvio_config_default(ctypes.byref(cfg))
    assert (cfg.levels, cfg.grid_size, cfg.max_iterations) == (6, 30, 30)
    h = ctypes.c_void_p()
    assert L.rsvio_tracker_create(None, ctypes.byref(h)) == -1
    cfg.width, cfg.height = 4, 4
    assert L.rsvio_tracker_create(ctypes.byref(cfg), ctypes.byref(h)) == -1
    assert L.rsvio_tracker_sync(None) == -1
    assert L.rsvio_status_string(-2).decode().startswith("CUDA")
    assert L.rsvio_pyramid_bytes(752, 480, 3) == 752 * 480 + 376 * 240 + 188 * 120


def test_no_cpu_fallback_without_gpu(rsvio):
    if rsvio.device_count() > 0:
        pytest.skip("a GPU is visible")
    with pytest.raises(rsvio.RsvioError):
        rsvio.StereoPatchTracker(40, 20, 0.01, width=752, height=480, levels=3)
    import numpy as np
    with pytest.raises(rsvio.RsvioError):
        rsvio.build_pyramid(np.zeros((64, 64), np.uint8), 3)


def test_rust_ffi_declares_every_header_symbol():
    """the Rust crate cannot be compiled here (no cargo); at least its raw bindings must name every entry point
    of include/rsvio_b200.h and nothing else"""
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    hdr = open(os.path.join(root, "include", "rsvio_b200.h")).read()
    ffi = open(os.path.join(root, "rs-vio_b200", "rust", "feature_tracker", "src", "ffi.rs")).read()
    in_header = set(_declared_functions())
    in_ffi = set(re.findall(r"pub fn (rsvio_[a-z0-9_]+)", ffi))
    assert in_header == in_ffi, (sorted(in_header - in_ffi), sorted(in_ffi - in_header))


def test_rust_build_script_lists_every_cuda_source():
    """build.rs (--features build-cuda) must compile the same .cu files as the Makefile"""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    mk = open(os.path.join(root, "rs-vio_b200", "Makefile")).read()
    srcs = set(re.findall(r"csrc/(\w+\.cu)", re.search(r"^SRCS\s*:=\s*(.*)$", mk, re.M).group(1)))
    rs = open(os.path.join(root, "rs-vio_b200", "rust", "feature_tracker", "build.rs")).read()
    assert srcs and srcs == set(re.findall(r'"(\w+\.cu)"', rs))


def test_tools_and_entry_points_compile():
    """bench.py, __graft_entry__.py and every script under tools/ at least parse (they only run on a GPU box)"""
    import py_compile
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    files = [os.path.join(root, "bench.py"), os.path.join(root, "__graft_entry__.py")]
    for d, _, names in os.walk(os.path.join(root, "tools")):
        files += [os.path.join(d, n) for n in names if n.endswith(".py")]
    assert len(files) > 8
    for f in files:
        py_compile.compile(f, doraise=True)


def test_crate_b_invalid_arguments_are_rejected_without_touching_the_gpu(rsvio):
    """row f-2 entry points: argument errors come back as RSVIO_ERR_INVALID_ARG (-1) before any CUDA call"""
    import numpy as np
    from rsvio_b200.binding import BConfig
    L = rsvio.load_library()
    cfg = BConfig()
    L.rsvio_b_config_default(ctypes.byref(cfg))
    assert (cfg.nlevels, cfg.ratio, cfg.preprocessing_blur, cfg.detection_min_dist, cfg.optical_flow_max_iter) == (5, 2.0, 1, 15, 25)
    assert abs(cfg.detection_threshold - 2.5) < 1e-6 and abs(cfg.detection_blur - 6.0) < 1e-6 and cfg.matching_cost == 0
    h = ctypes.c_void_p()
    assert L.rsvio_b_tracker_create(None, ctypes.byref(h)) == -1
    assert L.rsvio_b_tracker_create(ctypes.byref(cfg), ctypes.byref(h)) == -1            # width / height still 0
    cfg.width, cfg.height = 64, 64
    for field, bad in (("nlevels", 0), ("nlevels", 99), ("ratio", 1.0), ("detection_min_dist", 0), ("detection_blur", 0.0),
                       ("matching_cost", 7), ("optical_flow_max_iter", -1), ("max_features", -5)):
        c2 = BConfig.from_buffer_copy(cfg)
        setattr(c2, field, bad)
        assert L.rsvio_b_tracker_create(ctypes.byref(c2), ctypes.byref(h)) == -1, field
    img = np.zeros((8, 8), np.float32)
    assert L.rsvio_b_tracker_process_frame(None, img, 0) == -1
    n = ctypes.c_int32(0)
    assert L.rsvio_b_tracker_num_features(None, ctypes.byref(n), None) == -1
    assert L.rsvio_b_blur(0, img, 0, 8, 2.0, img) == -1
    assert L.rsvio_b_blur(0, img, 8, 8, 1000.0, img) == -1                             # kernel table holds sigma <= ~31
    assert L.rsvio_b_shi_tomasi_score(0, img, 8, 8, 0.0, img) == -1
    out = np.zeros((4, 2), np.uint32)
    cnt = ctypes.c_size_t(0)
    assert L.rsvio_b_add_points(0, img, 8, 8, img, 0, 0.0, 0, 2.0, out, 4, ctypes.byref(cnt)) == -1   # min_dist 0
    L.rsvio_b_tracker_destroy(None)                                                    # a no-op
