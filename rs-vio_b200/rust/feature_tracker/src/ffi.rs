//! Raw bindings of include/rsvio_b200.h (hand-written; keep in sync with RSVIO_ABI_VERSION = 1).
#![allow(non_camel_case_types)]
use std::os::raw::{c_char, c_void};

pub type rsvio_status = i32;
pub const RSVIO_OK: rsvio_status = 0;
pub const RSVIO_ERR_CAPACITY: rsvio_status = -3;
pub const RSVIO_MODE_STEREO: i32 = 0;
pub const RSVIO_MODE_MONO: i32 = 1;

#[repr(C)]
#[derive(Clone, Copy)]
pub struct rsvio_config {
    pub width: i32,
    pub height: i32,
    pub levels: i32,
    pub grid_size: u32,
    pub max_iterations: i32,
    pub convergence_threshold: f32,
    pub mode: i32,
    pub n_streams: i32,
    pub device: i32,
    pub max_tracks: i32,
    pub cuda_stream: *mut c_void,
}

#[repr(C)]
#[derive(Clone, Copy)]
pub struct rsvio_track_record {
    pub id: u64,
    pub x: f32,
    pub y: f32,
}

pub const RSVIO_PIXEL_LUMA8: i32 = 0;
pub const RSVIO_PIXEL_LUMA16: i32 = 1;
pub const RSVIO_PIXEL_RGB8: i32 = 2;
pub const RSVIO_CAMERA_NONE: i32 = 0;
pub const RSVIO_CAMERA_PINHOLE_RADTAN: i32 = 1;
pub const RSVIO_CAMERA_EUCM: i32 = 2;

#[repr(C)]
#[derive(Clone, Copy)]
pub struct rsvio_camera_model {
    pub model: i32,
    pub reserved: i32,
    pub params: [f64; 9],
}

#[repr(C)]
pub struct rsvio_tracker {
    _private: [u8; 0],
}

extern "C" {
    pub fn rsvio_abi_version() -> i32;
    pub fn rsvio_status_string(s: rsvio_status) -> *const c_char;
    pub fn rsvio_config_default(cfg: *mut rsvio_config);
    pub fn rsvio_host_alloc(ptr: *mut *mut c_void, bytes: usize) -> rsvio_status;
    pub fn rsvio_host_alloc_flags(ptr: *mut *mut c_void, bytes: usize, flags: u32) -> rsvio_status;
    pub fn rsvio_host_free(ptr: *mut c_void) -> rsvio_status;
    pub fn rsvio_tracker_create(cfg: *const rsvio_config, out: *mut *mut rsvio_tracker) -> rsvio_status;
    pub fn rsvio_tracker_destroy(t: *mut rsvio_tracker);
    pub fn rsvio_tracker_last_error(t: *const rsvio_tracker) -> *const c_char;
    pub fn rsvio_tracker_reset(t: *mut rsvio_tracker) -> rsvio_status;
    pub fn rsvio_tracker_process_stereo(t: *mut rsvio_tracker, left: *const u8, right: *const u8, stride: usize) -> rsvio_status;
    pub fn rsvio_tracker_process_mono(t: *mut rsvio_tracker, image: *const u8, stride: usize) -> rsvio_status;
    pub fn rsvio_tracker_result_view(t: *mut rsvio_tracker, records: *mut *const rsvio_track_record, counts: *mut *const i32,
                                     groups: *mut i32, capacity: *mut i32) -> rsvio_status;
    pub fn rsvio_tracker_remove_ids(t: *mut rsvio_tracker, stream: i32, ids: *const u64, n: usize) -> rsvio_status;

    // ---- batched / asynchronous forms (n_streams > 1; at most two frames in flight) ----
    pub fn rsvio_device_count(n: *mut i32) -> rsvio_status;
    pub fn rsvio_tracker_process_batch(t: *mut rsvio_tracker, left: *const *const u8, right: *const *const u8,
                                       stride: usize) -> rsvio_status;
    pub fn rsvio_tracker_submit_batch(t: *mut rsvio_tracker, left: *const *const u8, right: *const *const u8,
                                      stride: usize) -> rsvio_status;
    pub fn rsvio_tracker_submit_device(t: *mut rsvio_tracker, d_left: *const c_void, d_right: *const c_void,
                                       stride: usize, pitch: usize) -> rsvio_status;
    pub fn rsvio_tracker_sync(t: *mut rsvio_tracker) -> rsvio_status;

    // ---- copies of the track tables ----
    pub fn rsvio_tracker_num_tracks(t: *mut rsvio_tracker, stream: i32, cam: i32, n: *mut i32) -> rsvio_status;
    pub fn rsvio_tracker_get_tracks(t: *mut rsvio_tracker, stream: i32, cam: i32, ids: *mut u64, xy: *mut f32,
                                    cap: usize, n: *mut usize) -> rsvio_status;
    pub fn rsvio_tracker_get_transforms(t: *mut rsvio_tracker, stream: i32, cam: i32, ids: *mut u64, t6: *mut f32,
                                        cap: usize, n: *mut usize) -> rsvio_status;

    // ---- counters and per-stage device times ----
    pub fn rsvio_tracker_counters(t: *mut rsvio_tracker, out: *mut u64) -> rsvio_status; // out[6]
    pub fn rsvio_tracker_overflow_frames(t: *mut rsvio_tracker, n: *mut u64) -> rsvio_status;
    pub fn rsvio_tracker_set_profiling(t: *mut rsvio_tracker, enabled: i32) -> rsvio_status;
    pub fn rsvio_tracker_stage_times(t: *mut rsvio_tracker, ms: *mut f32, launches: *mut u64, reset: i32)
        -> rsvio_status; // ms[RSVIO_N_STAGES], launches[RSVIO_N_STAGES]

    // ---- stateless entry points (parity tests, kernel benchmarks) ----
    pub fn rsvio_pyramid_bytes(width: i32, height: i32, levels: i32) -> usize;
    pub fn rsvio_pyramid_device_bytes(width: i32, height: i32, levels: i32) -> usize;
    pub fn rsvio_build_pyramid_u8(device: i32, images: *const u8, n_images: i32, width: i32, height: i32,
                                  stride: usize, pitch: usize, levels: i32, out: *mut u8) -> rsvio_status;
    pub fn rsvio_build_pyramid_u8_device(device: i32, d_images: *const c_void, n_images: i32, width: i32, height: i32,
                                         stride: usize, pitch: usize, levels: i32, d_out: *mut c_void, repeats: i32,
                                         kernel_ms: *mut f32) -> rsvio_status;
    pub fn rsvio_track_points(device: i32, pyr_a: *const u8, pyr_b: *const u8, width: i32, height: i32, levels: i32,
                              n: i32, in_t6: *const f32, max_iterations: i32, convergence_threshold: f32,
                              out_t6: *mut f32, status: *mut i32, iters: *mut i32) -> rsvio_status;
    pub fn rsvio_detect_keypoints(device: i32, image: *const u8, width: i32, height: i32, stride: usize,
                                  grid_size: u32, cur_xy: *const u32, n_cur: i32, out_xy: *mut u32,
                                  out_score: *mut f32, cap: usize, n: *mut usize) -> rsvio_status;
    pub fn rsvio_microbench_fp32(device: i32, ffma_tflops: *mut f32, fmul_fadd_tflops: *mut f32) -> rsvio_status;
    pub fn rsvio_se2_exp(device: i32, twists: *const f32, n: i32, out9: *mut f32) -> rsvio_status;

    // ---- unprojection of the tracked pixels (Feature::undistorted_coord) ----
    pub fn rsvio_tracker_set_camera(t: *mut rsvio_tracker, cam: i32, model: *const rsvio_camera_model) -> rsvio_status;
    pub fn rsvio_tracker_undistorted_view(t: *mut rsvio_tracker, uxy: *mut *const f32, groups: *mut i32,
                                          capacity: *mut i32) -> rsvio_status;
    pub fn rsvio_unproject_points(device: i32, model: *const rsvio_camera_model, xy: *const f32, n: i32,
                                  out_xy: *mut f32) -> rsvio_status;

    // ---- decoded frames (16-bit grey, RGB8): DynamicImage::to_luma8 on the device ----
    pub fn rsvio_tracker_process_stereo_fmt(t: *mut rsvio_tracker, left: *const c_void, right: *const c_void,
                                            stride_bytes: usize, pixel_format: i32) -> rsvio_status;
    pub fn rsvio_convert_to_luma8(device: i32, src: *const c_void, width: i32, height: i32, stride_bytes: usize,
                                  pixel_format: i32, out: *mut u8) -> rsvio_status;

    // ---- row f-2, first slice: crate B's track_points between two f32 pyramids ----
    pub fn rsvio_b_pyramid_dims(width: i32, height: i32, levels: i32, ratio: f64, widths: *mut i32, heights: *mut i32)
        -> rsvio_status;
    pub fn rsvio_b_build_pyramid(device: i32, image: *const f32, width: i32, height: i32, levels: i32, ratio: f64,
                                 out_levels: *const *mut f32) -> rsvio_status;
    pub fn rsvio_b_track_points(device: i32, pyr0: *const *const f32, pyr1: *const *const f32, widths: *const i32,
                                heights: *const i32, levels: i32, points_xy: *const f32, n: i32, max_iteration: i32,
                                lm_lambda: f32, out_t4: *mut f32, status: *mut i32) -> rsvio_status;

    // ---- row f-2: the rest of crate B (blur, LSSD, Shi-Tomasi detection, stateful FeatureTracker) ----
    pub fn rsvio_b_track_points_cost(device: i32, pyr0: *const *const f32, pyr1: *const *const f32, widths: *const i32,
                                     heights: *const i32, levels: i32, points_xy: *const f32, n: i32,
                                     max_iteration: i32, lm_lambda: f32, matching_cost: i32, out_t4: *mut f32,
                                     status: *mut i32) -> rsvio_status;
    pub fn rsvio_b_blur(device: i32, image: *const f32, width: i32, height: i32, sigma: f32, out: *mut f32) -> rsvio_status;
    pub fn rsvio_b_shi_tomasi_score(device: i32, image: *const f32, width: i32, height: i32, detection_blur: f32,
                                    out_score: *mut f32) -> rsvio_status;
    pub fn rsvio_b_add_points(device: i32, image: *const f32, width: i32, height: i32, tracked_xy: *const f32,
                              n_tracked: i32, threshold: f32, min_dist: u32, detection_blur: f32, out_xy: *mut u32,
                              cap: usize, n: *mut usize) -> rsvio_status;
    pub fn rsvio_b_config_default(cfg: *mut rsvio_b_config);
    pub fn rsvio_b_tracker_create(cfg: *const rsvio_b_config, out: *mut *mut rsvio_b_tracker) -> rsvio_status;
    pub fn rsvio_b_tracker_destroy(t: *mut rsvio_b_tracker);
    pub fn rsvio_b_tracker_process_frame(t: *mut rsvio_b_tracker, image: *const f32, frame_id: i64) -> rsvio_status;
    pub fn rsvio_b_tracker_num_features(t: *const rsvio_b_tracker, n: *mut i32, n_new: *mut i32) -> rsvio_status;
    pub fn rsvio_b_tracker_get_features(t: *const rsvio_b_tracker, ids: *mut u64, xy: *mut f32, cap: usize,
                                        n: *mut usize) -> rsvio_status;
    pub fn rsvio_b_tracker_pyramid_dims(t: *const rsvio_b_tracker, levels: *mut i32, widths: *mut i32,
                                        heights: *mut i32) -> rsvio_status;
    pub fn rsvio_b_tracker_get_pyramid_level(t: *mut rsvio_b_tracker, level: i32, out: *mut f32) -> rsvio_status;
    pub fn rsvio_b_tracker_last_frame_ms(t: *const rsvio_b_tracker, ms: *mut f32) -> rsvio_status;
}

pub const RSVIO_B_COST_SSD: i32 = 0;
pub const RSVIO_B_COST_LSSD: i32 = 1;

/// rsvio_b_config (include/rsvio_b200.h): crate B's FeatureTrackingConfig + image size, device, matching cost
#[repr(C)]
#[derive(Clone, Copy, Debug)]
pub struct rsvio_b_config {
    pub width: i32,
    pub height: i32,
    pub nlevels: i32,
    pub ratio: f64,
    pub preprocessing_blur: i32,
    pub preprocessing_blur_sigma: f32,
    pub detection_threshold: f32,
    pub detection_min_dist: u32,
    pub detection_blur: f32,
    pub optical_flow_max_iter: i32,
    pub optical_flow_lm_lambda: f32,
    pub matching_cost: i32,
    pub device: i32,
    pub max_features: i32,
}

#[repr(C)]
pub struct rsvio_b_tracker {
    _private: [u8; 0],
}

pub const RSVIO_N_STAGES: usize = 8;
