"""rsvio_b200 -- Python host side of the B200-native stereo patch tracker.

Thin ctypes layer over the C ABI in include/rsvio_b200.h (librsvio_b200.so, CUDA sm_100a) that
mirrors the reference's tracker API (src/feature_tracker/feature_tracker.rs): StereoPatchTracker,
PatchTracker, Feature.  There is no CPU fallback: loading fails loudly when the library is missing
and every call fails with RsvioError when no CUDA device is usable.
"""
from .binding import (  # noqa: F401
    ABI_SYMBOLS, CAMERA_EUCM, CAMERA_NONE, CAMERA_PINHOLE_RADTAN, PIXEL_LUMA8, PIXEL_LUMA16, PIXEL_RGB8, CameraModel, CapacityWarning,
    Feature, HostBuffer, PatchTracker, RsvioError, b_build_pyramid, b_track_points, convert_to_luma8,
    B_COST_LSSD, B_COST_SSD, BConfig, FeatureTrackerB, b_add_points, b_blur, b_shi_tomasi_score, b_track_points_cost,
    StereoPatchTracker, build_library, build_pyramid, detect_keypoints, device_count, library_path, load_library, microbench_fp32,
    pyramid_bytes, se2_exp, split_pyramid, track_points, unproject_points,
)
from . import sharding, synth  # noqa: F401
