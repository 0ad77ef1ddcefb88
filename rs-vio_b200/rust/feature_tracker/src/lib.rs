//! GPU drop-in for `rs_vio::feature_tracker` (reference src/feature_tracker/feature_tracker.rs).
//!
//! Same public items and signatures as the reference module so that `src/estimator/estimator.rs:27,85-89,190`
//! compiles unchanged after replacing `use crate::feature_tracker::...` by this crate:
//!   Feature::new                                  feature_tracker.rs:25
//!   StereoPatchTracker::<N>::new / process_frame / get_track_points / remove_id   :103 / :116 / :188 / :201
//!   PatchTracker::<N>::process_frame / get_track_points / remove_id               :43 / :78 / :84
//! The tracker owns an opaque handle of librsvio_b200.so; all kernels run on the B200.  There is no
//! CPU fallback: construction panics with the library's error text when no CUDA device is usable
//! (the reference's signatures are infallible, so there is no Result to put it in).
//!
//! NOT COMPILED IN THIS REPOSITORY (no Rust toolchain in the image) -- INTEGRATION.md.
pub mod ffi;
/// the reference's standalone `feature_tracker` crate (mono, f32 pyramid, bicubic, Shi-Tomasi): row f-2
pub mod crate_b;

use image::GrayImage;
use std::collections::HashMap;
use std::ffi::CStr;

#[derive(Debug, Clone)]
pub struct Feature {
    pub feature_id: usize,
    pub pixel_coord: [f32; 2],
    /// `[-1, -1]` means invalid; filled by `Frame::add_left_feature` / `add_right_feature`.
    pub undistorted_coord: [f32; 2],
}

impl Feature {
    pub fn new(feature_id: usize, pixel_coord: [f32; 2]) -> Self {
        Self { feature_id, pixel_coord, undistorted_coord: [-1.0, -1.0] }
    }
}

/// What `process_frame` needs from `crate::estimator::Frame` (src/estimator/frame.rs:107-134).
pub trait FeatureSink {
    fn add_left_feature(&mut self, f: Feature);
    fn add_right_feature(&mut self, f: Feature);
}

struct Handle {
    ptr: *mut ffi::rsvio_tracker,
    dims: (u32, u32),
    cfg: ffi::rsvio_config,
    cameras: [Option<ffi::rsvio_camera_model>; 2],
}
// one handle == one CUDA stream; `&mut self` on every call gives the exclusivity the C ABI asks for
unsafe impl Send for Handle {}

impl Handle {
    fn new(levels: u32, grid: u32, max_it: u32, thr: f32, mode: i32) -> Self {
        let mut cfg = unsafe { std::mem::zeroed::<ffi::rsvio_config>() };
        unsafe { ffi::rsvio_config_default(&mut cfg) };
        cfg.levels = levels as i32;
        cfg.grid_size = grid;
        cfg.max_iterations = max_it as i32;
        cfg.convergence_threshold = thr;
        cfg.mode = mode;
        Self { ptr: std::ptr::null_mut(), dims: (0, 0), cfg, cameras: [None, None] }
    }
    /// The reference learns the image size from the first frame; so does the handle.
    fn ensure(&mut self, w: u32, h: u32) {
        if !self.ptr.is_null() && self.dims == (w, h) {
            return;
        }
        self.release();
        let mut cfg = self.cfg;
        cfg.width = w as i32;
        cfg.height = h as i32;
        let st = unsafe { ffi::rsvio_tracker_create(&cfg, &mut self.ptr) };
        assert!(st == ffi::RSVIO_OK, "rsvio_tracker_create: {}", status_text(st));
        self.dims = (w, h);
        for cam in 0..2 {
            if let Some(model) = self.cameras[cam] {
                let st = unsafe { ffi::rsvio_tracker_set_camera(self.ptr, cam as i32, &model) };
                self.check(st, "rsvio_tracker_set_camera");
            }
        }
    }
    fn check(&self, st: ffi::rsvio_status, what: &str) {
        if st != ffi::RSVIO_OK {
            let detail = unsafe { CStr::from_ptr(ffi::rsvio_tracker_last_error(self.ptr)) }.to_string_lossy().into_owned();
            panic!("{what}: {} ({detail})", status_text(st));
        }
    }
    /// Status of a `process_*` call.  A full track table is not an error: the frame was processed and is delivered,
    /// only new points were dropped (the reference's maps are unbounded and its `process_frame` cannot fail,
    /// feature_tracker.rs:93-96,116).  One warning on stderr per handle; the count stays queryable
    /// (`rsvio_tracker_overflow_frames`).
    fn check_frame(&self, st: ffi::rsvio_status, what: &str) {
        if st == ffi::RSVIO_ERR_CAPACITY {
            let mut n: u64 = 0;
            if unsafe { ffi::rsvio_tracker_overflow_frames(self.ptr, &mut n) } == ffi::RSVIO_OK && n == 1 {
                let detail = unsafe { CStr::from_ptr(ffi::rsvio_tracker_last_error(self.ptr)) }.to_string_lossy().into_owned();
                eprintln!("rsvio: {what}: {detail}");
            }
            return;
        }
        self.check(st, what);
    }
    /// Camera model of one camera (0 = left, 1 = right); applied when the handle exists (the reference builds its
    /// models from the YAML config, src/datasets/mod.rs:93-163).  With a model set, `process_frame` hands out
    /// features whose `undistorted_coord` is already filled (what `Frame::add_*_feature` computes on the CPU).
    fn set_camera(&mut self, cam: usize, model: ffi::rsvio_camera_model) {
        self.cameras[cam] = Some(model);
        if !self.ptr.is_null() {
            let st = unsafe { ffi::rsvio_tracker_set_camera(self.ptr, cam as i32, &model) };
            self.check(st, "rsvio_tracker_set_camera");
        }
    }
    /// zero-copy view of the host-mapped result the export kernel wrote; `u` is the unprojected coordinate or
    /// `[-1, -1]` (the reference's placeholder) when no camera model is set
    fn for_each_feature(&self, cam: usize, mut f: impl FnMut(Feature)) {
        if self.ptr.is_null() {
            return;
        }
        let mut uxy: *const f32 = std::ptr::null();
        let (mut g2, mut c2) = (0i32, 0i32);
        let have_u = unsafe { ffi::rsvio_tracker_undistorted_view(self.ptr, &mut uxy, &mut g2, &mut c2) } == ffi::RSVIO_OK
            && !uxy.is_null();
        let cap = c2 as usize;
        let groups = g2 as usize;
        let mut k = 0usize;
        self.for_each(cam, |id, x, y| {
            let mut feat = Feature::new(id, [x, y]);
            if have_u {
                let base = (cam * (groups / 2) * cap + k) * 2;
                feat.undistorted_coord = unsafe { [*uxy.add(base), *uxy.add(base + 1)] };
            }
            k += 1;
            f(feat);
        });
    }
    /// zero-copy view of the host-mapped result the export kernel wrote
    fn for_each(&self, cam: usize, mut f: impl FnMut(usize, f32, f32)) {
        if self.ptr.is_null() {
            return;
        }
        let (mut rec, mut cnt) = (std::ptr::null(), std::ptr::null());
        let (mut groups, mut cap) = (0i32, 0i32);
        let st = unsafe { ffi::rsvio_tracker_result_view(self.ptr, &mut rec, &mut cnt, &mut groups, &mut cap) };
        if st != ffi::RSVIO_OK {
            return; // no frame processed yet
        }
        let g = cam * (groups as usize / 2);
        let n = unsafe { *cnt.add(g) } as usize;
        let base = unsafe { rec.add(g * cap as usize) };
        for k in 0..n {
            let r = unsafe { *base.add(k) };
            f(r.id as usize, r.x, r.y);
        }
    }
    fn remove(&mut self, ids: &[usize]) {
        if self.ptr.is_null() || ids.is_empty() {
            return;
        }
        let v: Vec<u64> = ids.iter().map(|&i| i as u64).collect();
        let st = unsafe { ffi::rsvio_tracker_remove_ids(self.ptr, 0, v.as_ptr(), v.len()) };
        self.check(st, "rsvio_tracker_remove_ids");
    }
    fn release(&mut self) {
        if !self.ptr.is_null() {
            unsafe { ffi::rsvio_tracker_destroy(self.ptr) };
            self.ptr = std::ptr::null_mut();
        }
    }
}

impl Drop for Handle {
    fn drop(&mut self) {
        self.release();
    }
}

pub(crate) fn status_text(st: ffi::rsvio_status) -> String {
    unsafe { CStr::from_ptr(ffi::rsvio_status_string(st)) }.to_string_lossy().into_owned()
}

pub struct StereoPatchTracker<const N: u32> {
    h: Handle,
}

impl<const LEVELS: u32> StereoPatchTracker<LEVELS> {
    pub fn new(grid_size: u32, optical_flow_max_iterations: u32, optical_flow_convergence_threshold: f64) -> Self {
        Self {
            h: Handle::new(LEVELS, grid_size, optical_flow_max_iterations, optical_flow_convergence_threshold as f32,
                           ffi::RSVIO_MODE_STEREO),
        }
    }

    pub fn process_frame<F: FeatureSink>(&mut self, greyscale_image0: &GrayImage, greyscale_image1: &GrayImage, frame: &mut F) {
        let (w, h) = greyscale_image0.dimensions();
        assert_eq!((w, h), greyscale_image1.dimensions(), "stereo images differ in size");
        self.h.ensure(w, h);
        let st = unsafe {
            ffi::rsvio_tracker_process_stereo(self.h.ptr, greyscale_image0.as_raw().as_ptr(), greyscale_image1.as_raw().as_ptr(),
                                              w as usize)
        };
        self.h.check_frame(st, "rsvio_tracker_process_stereo");
        self.h.for_each_feature(0, |f| frame.add_left_feature(f));
        self.h.for_each_feature(1, |f| frame.add_right_feature(f));
    }

    /// Optional: pinhole-radtan parameters `[fx, fy, cx, cy, k1, k2, p1, p2, k3]` of one camera (0 = left,
    /// 1 = right).  The features handed to `frame` then carry `undistorted_coord` and `Frame::add_*_feature`
    /// can skip its own `unproject_one` (src/estimator/frame.rs:118,131).
    pub fn set_pinhole_radtan(&mut self, cam: usize, params: [f64; 9]) {
        self.h.set_camera(cam, ffi::rsvio_camera_model { model: ffi::RSVIO_CAMERA_PINHOLE_RADTAN, reserved: 0, params });
    }

    /// Optional: EUCM parameters `[fx, fy, cx, cy, alpha, beta]` of one camera.
    pub fn set_eucm(&mut self, cam: usize, p: [f64; 6]) {
        let params = [p[0], p[1], p[2], p[3], p[4], p[5], 0.0, 0.0, 0.0];
        self.h.set_camera(cam, ffi::rsvio_camera_model { model: ffi::RSVIO_CAMERA_EUCM, reserved: 0, params });
    }

    pub fn get_track_points(&self) -> [HashMap<usize, (f32, f32)>; 2] {
        let mut out = [HashMap::new(), HashMap::new()];
        for cam in 0..2 {
            self.h.for_each(cam, |id, x, y| {
                out[cam].insert(id, (x, y));
            });
        }
        out
    }

    pub fn remove_id(&mut self, ids: &[usize]) {
        self.h.remove(ids);
    }
}

pub struct PatchTracker<const N: u32> {
    h: Handle,
}

impl<const LEVELS: u32> Default for PatchTracker<LEVELS> {
    fn default() -> Self {
        // DEFAULT_GRID_SIZE 30, 30 iterations, 0.005 (feature_tracker.rs:51-52,65)
        Self { h: Handle::new(LEVELS, 30, 30, 0.005, ffi::RSVIO_MODE_MONO) }
    }
}

impl<const LEVELS: u32> PatchTracker<LEVELS> {
    pub fn process_frame(&mut self, greyscale_image: &GrayImage) {
        let (w, h) = greyscale_image.dimensions();
        self.h.ensure(w, h);
        let st = unsafe { ffi::rsvio_tracker_process_mono(self.h.ptr, greyscale_image.as_raw().as_ptr(), w as usize) };
        self.h.check_frame(st, "rsvio_tracker_process_mono");
    }
    pub fn get_track_points(&self) -> HashMap<usize, (f32, f32)> {
        let mut out = HashMap::new();
        self.h.for_each(0, |id, x, y| {
            out.insert(id, (x, y));
        });
        out
    }
    pub fn remove_id(&mut self, ids: &[usize]) {
        self.h.remove(ids);
    }
}

/// The reference's `image_utilities` module (src/feature_tracker/image_utilities.rs, `pub mod` at mod.rs:1) with the
/// types of this crate: plain predicates on the host, `se2_exp_matrix` and `detect_key_points` evaluated by the device
/// code through the stateless entry points (the same kernels the tracker runs).  `image_grad`, `Pattern52` and
/// `track_point_at_level` (mod.rs:5, feature_tracker.rs:344) are internal to the fused LK kernel and have no
/// standalone entry point: `rsvio_track_points` (ffi) runs the whole coarse-to-fine chain for a batch of points.
pub mod image_utilities {
    use crate::ffi;
    use image::GrayImage;

    /// imageproc::corners::Corner as the reference uses it (x, y, score)
    #[derive(Copy, Clone, Debug, PartialEq)]
    pub struct Corner {
        pub x: u32,
        pub y: u32,
        pub score: f32,
    }

    /// image_utilities.rs:68-73
    pub fn point_in_bound(keypoint: &Corner, height: u32, width: u32, radius: u32) -> bool {
        keypoint.x >= radius && keypoint.x <= width - radius && keypoint.y >= radius && keypoint.y <= height - radius
    }

    /// image_utilities.rs:75-80 (`as u32` saturates: negatives and NaN become 0)
    pub fn inbound(image: &GrayImage, x: f32, y: f32, radius: u32) -> bool {
        let x = x.round() as u32;
        let y = y.round() as u32;
        x >= radius && y >= radius && x < image.width() - radius && y < image.height() - radius
    }

    /// image_utilities.rs:82-106: twist `[vx, vy, theta]` -> row-major 3x3 homogeneous matrix
    pub fn se2_exp_matrix(a: &[f32; 3]) -> [[f32; 3]; 3] {
        let mut out = [0f32; 9];
        let st = unsafe { ffi::rsvio_se2_exp(0, a.as_ptr(), 1, out.as_mut_ptr()) };
        assert!(st == ffi::RSVIO_OK, "rsvio_se2_exp: {}", crate::status_text(st));
        [[out[0], out[1], out[2]], [out[3], out[4], out[5]], [out[6], out[7], out[8]]]
    }

    /// image_utilities.rs:108-175 with `num_points_in_cell == 1` (the only value the reference passes,
    /// feature_tracker.rs:242): one FAST-9 corner per empty grid cell, in the reference's visiting order
    pub fn detect_key_points(image: &GrayImage, grid_size: u32, current_corners: &Vec<Corner>, num_points_in_cell: u32) -> Vec<Corner> {
        assert_eq!(num_points_in_cell, 1, "the device path implements one point per cell");
        let (w, h) = image.dimensions();
        let cur: Vec<u32> = current_corners.iter().flat_map(|c| [c.x, c.y]).collect();
        let cap = ((w / grid_size + 1) * (h / grid_size + 1)) as usize;
        let mut xy = vec![0u32; 2 * cap];
        let mut score = vec![0f32; cap];
        let mut n = 0usize;
        let cur_ptr = if cur.is_empty() { std::ptr::null() } else { cur.as_ptr() };
        let st = unsafe {
            ffi::rsvio_detect_keypoints(0, image.as_raw().as_ptr(), w as i32, h as i32, w as usize, grid_size, cur_ptr,
                                        current_corners.len() as i32, xy.as_mut_ptr(), score.as_mut_ptr(), cap, &mut n)
        };
        assert!(st == ffi::RSVIO_OK, "rsvio_detect_keypoints: {}", crate::status_text(st));
        (0..n.min(cap)).map(|i| Corner { x: xy[2 * i], y: xy[2 * i + 1], score: score[i] }).collect()
    }
}
