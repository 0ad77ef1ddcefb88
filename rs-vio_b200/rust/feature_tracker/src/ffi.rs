//! Raw bindings of include/rsvio_b200.h (hand-written; keep in sync with RSVIO_ABI_VERSION = 1).
#![allow(non_camel_case_types)]
use std::os::raw::{c_char, c_void};

pub type rsvio_status = i32;
pub const RSVIO_OK: rsvio_status = 0;
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

#[repr(C)]
pub struct rsvio_tracker {
    _private: [u8; 0],
}

extern "C" {
    pub fn rsvio_abi_version() -> i32;
    pub fn rsvio_status_string(s: rsvio_status) -> *const c_char;
    pub fn rsvio_config_default(cfg: *mut rsvio_config);
    pub fn rsvio_host_alloc(ptr: *mut *mut c_void, bytes: usize) -> rsvio_status;
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
}
