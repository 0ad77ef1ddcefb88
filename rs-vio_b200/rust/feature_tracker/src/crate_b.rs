//! GPU drop-in for the reference's standalone `feature_tracker` crate ("crate B", row f-2):
//!   FeatureTrackingConfig                        feature_tracker/src/feature_tracker.rs:24-38
//!   Feature { feature_id, central_point }        :41-48
//!   FeatureTracker::new / process_frame / get_pyramid   :59-69 / :77-185 / :190-193
//!   ext::Frame { frame_id, features }            feature_tracker/src/ext.rs:3-15
//!   patch::MatchingCost                          feature_tracker/src/patch.rs:4-8
//! Same names and signatures; the viewer parameter of `new` is accepted and ignored (the reference only logs through
//! it).  `process_frame` keeps the frame it is given, as the reference does (`previous_frame_pyramid`), and exposes
//! it through `last_frame()`; `get_pyramid()` downloads the device pyramid on demand.
//!
//! NOT COMPILED IN THIS REPOSITORY (no Rust toolchain in the image) -- INTEGRATION.md.
use crate::ffi;
use image::{ImageBuffer, Luma};
use nalgebra as na;
use serde::{Deserialize, Serialize};

pub type Float = f32;
pub type FloatGrayImage = ImageBuffer<Luma<Float>, Vec<Float>>;
pub type Pyramid = Vec<FloatGrayImage>;

#[derive(Copy, Clone, Debug)]
pub enum MatchingCost {
    SSD,
    LSSD,
}

#[derive(Serialize, Deserialize, Debug)]
pub struct FeatureTrackingConfig {
    pub nlevels: usize,
    pub ratio: f64,
    pub preprocessing_blur: bool,
    pub preprocessing_blur_sigma: Float,
    pub detection_threshold: Float,
    pub detection_min_dist: u32,
    pub detection_blur: Float,
    pub optical_flow_max_iter: usize,
    pub optical_flow_lm_lambda: Float,
}

#[derive(Debug, Clone)]
pub struct Feature {
    pub feature_id: usize,
    pub central_point: na::Point2<Float>,
}

pub mod ext {
    use super::Feature;
    pub struct Frame {
        pub frame_id: i64,
        pub features: Vec<Feature>,
    }
    impl Frame {
        pub fn new(id: i64) -> Frame {
            Frame { frame_id: id, features: Vec::new() }
        }
    }
}

/// stands in for `&dyn FeatureTrackerViewer` (feature_tracker/src/viewer.rs): logging hooks only
pub trait FeatureTrackerViewer {}

pub struct FeatureTracker<'a> {
    config: FeatureTrackingConfig,
    matching_cost: MatchingCost,
    handle: *mut ffi::rsvio_b_tracker,
    dims: (u32, u32),
    previous_frame: Option<ext::Frame>,
    pyramid: Option<Pyramid>,
    _viewer: Option<&'a dyn FeatureTrackerViewer>,
}
unsafe impl<'a> Send for FeatureTracker<'a> {}

impl<'a> FeatureTracker<'a> {
    pub fn new(config: FeatureTrackingConfig, viewer: Option<&'a dyn FeatureTrackerViewer>) -> Self {
        FeatureTracker {
            config,
            matching_cost: MatchingCost::SSD, // process_frame hard-codes SSD (feature_tracker.rs:125)
            handle: std::ptr::null_mut(),
            dims: (0, 0),
            previous_frame: None,
            pyramid: None,
            _viewer: viewer,
        }
    }

    /// not in the reference: selects the cost its `patch` module also offers
    pub fn set_matching_cost(&mut self, cost: MatchingCost) {
        self.matching_cost = cost;
        self.release();
    }

    fn release(&mut self) {
        if !self.handle.is_null() {
            unsafe { ffi::rsvio_b_tracker_destroy(self.handle) };
            self.handle = std::ptr::null_mut();
        }
    }

    /// the handle learns the image size from the first frame, like the reference's pyramids
    fn ensure(&mut self, w: u32, h: u32) {
        if !self.handle.is_null() && self.dims == (w, h) {
            return;
        }
        self.release();
        let mut cfg = unsafe { std::mem::zeroed::<ffi::rsvio_b_config>() };
        unsafe { ffi::rsvio_b_config_default(&mut cfg) };
        cfg.width = w as i32;
        cfg.height = h as i32;
        cfg.nlevels = self.config.nlevels as i32;
        cfg.ratio = self.config.ratio;
        cfg.preprocessing_blur = self.config.preprocessing_blur as i32;
        cfg.preprocessing_blur_sigma = self.config.preprocessing_blur_sigma;
        cfg.detection_threshold = self.config.detection_threshold;
        cfg.detection_min_dist = self.config.detection_min_dist;
        cfg.detection_blur = self.config.detection_blur;
        cfg.optical_flow_max_iter = self.config.optical_flow_max_iter as i32;
        cfg.optical_flow_lm_lambda = self.config.optical_flow_lm_lambda;
        cfg.matching_cost = match self.matching_cost {
            MatchingCost::SSD => ffi::RSVIO_B_COST_SSD,
            MatchingCost::LSSD => ffi::RSVIO_B_COST_LSSD,
        };
        let st = unsafe { ffi::rsvio_b_tracker_create(&cfg, &mut self.handle) };
        assert!(st == ffi::RSVIO_OK, "rsvio_b_tracker_create: {}", crate::status_text(st));
        self.dims = (w, h);
    }

    pub fn process_frame(&mut self, in_image: &FloatGrayImage, mut frame: ext::Frame) {
        let (w, h) = in_image.dimensions();
        self.ensure(w, h);
        let st = unsafe { ffi::rsvio_b_tracker_process_frame(self.handle, in_image.as_raw().as_ptr(), frame.frame_id) };
        // a full feature table drops new corners, it does not fail the frame (the reference's Vec is unbounded)
        assert!(st == ffi::RSVIO_OK || st == ffi::RSVIO_ERR_CAPACITY, "rsvio_b_tracker_process_frame: {}", crate::status_text(st));
        let (mut n, mut n_new) = (0i32, 0i32);
        unsafe { ffi::rsvio_b_tracker_num_features(self.handle, &mut n, &mut n_new) };
        let mut ids = vec![0u64; n as usize];
        let mut xy = vec![0f32; 2 * n as usize];
        let mut got = 0usize;
        unsafe { ffi::rsvio_b_tracker_get_features(self.handle, ids.as_mut_ptr(), xy.as_mut_ptr(), n as usize, &mut got) };
        // `frame.features.extend(tracked_features)` then `.extend(new_features)` (feature_tracker.rs:144,173)
        frame.features.extend((0..n as usize).map(|i| Feature {
            feature_id: ids[i] as usize,
            central_point: na::Point2::new(xy[2 * i], xy[2 * i + 1]),
        }));
        self.previous_frame = Some(frame);
        self.pyramid = None;
    }

    /// the frame of the last `process_frame` call (the reference keeps it in `previous_frame_pyramid`)
    pub fn last_frame(&self) -> Option<&ext::Frame> {
        self.previous_frame.as_ref()
    }

    pub fn get_pyramid(&mut self) -> Option<&Pyramid> {
        if self.previous_frame.is_none() {
            return None;
        }
        if self.pyramid.is_none() {
            let mut levels = 0i32;
            let (mut ws, mut hs) = ([0i32; 8], [0i32; 8]);
            unsafe { ffi::rsvio_b_tracker_pyramid_dims(self.handle, &mut levels, ws.as_mut_ptr(), hs.as_mut_ptr()) };
            let mut pyr = Vec::new();
            for l in 0..levels as usize {
                let mut buf = vec![0f32; (ws[l] * hs[l]) as usize];
                let st = unsafe { ffi::rsvio_b_tracker_get_pyramid_level(self.handle, l as i32, buf.as_mut_ptr()) };
                assert!(st == ffi::RSVIO_OK, "rsvio_b_tracker_get_pyramid_level: {}", crate::status_text(st));
                pyr.push(FloatGrayImage::from_vec(ws[l] as u32, hs[l] as u32, buf).unwrap());
            }
            self.pyramid = Some(pyr);
        }
        self.pyramid.as_ref()
    }
}

impl<'a> Drop for FeatureTracker<'a> {
    fn drop(&mut self) {
        self.release();
    }
}
