// feature_tracker.hpp -- C++ host mirror of RS-VIO's tracker API over the C ABI (include/rsvio_b200.h).
//
// The reference is Rust (no toolchain in this image), so the host side above the C ABI is written in
// C++ with the reference's names, argument meaning and error behaviour:
//   rsvio::Feature                      src/feature_tracker/feature_tracker.rs:12-32
//   rsvio::StereoPatchTracker<LEVELS>   :91-207  (new :103, process_frame :116, get_track_points :188, remove_id :201)
//   rsvio::PatchTracker<LEVELS>         :36-89   (process_frame :43, get_track_points :78, remove_id :84)
// `Frame` is any type with add_left_feature(Feature) / add_right_feature(Feature)
// (src/estimator/frame.rs:107-134).  Infallible signatures like the reference: errors of the GPU
// layer surface as rsvio::Error exceptions at construction, and as std::runtime_error from the
// per-frame calls (a reference user never sees per-point failures either: points just disappear).
#pragma once

#include <cstdint>
#include <cstdio>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <utility>
#include <vector>

#include "../../../include/rsvio_b200.h"

namespace rsvio {

struct Error : std::runtime_error {
    rsvio_status status;
    Error(rsvio_status s, const std::string &what) : std::runtime_error(what), status(s) {}
};

struct Feature {
    std::size_t feature_id;
    float pixel_coord[2];
    float undistorted_coord[2];  // [-1, -1] == invalid, filled by Frame::add_*_feature
    Feature(std::size_t id, float x, float y) : feature_id(id), pixel_coord{x, y}, undistorted_coord{-1.0f, -1.0f} {}
};

// GrayImage: row-major 8-bit luma, stride == width (image::GrayImage)
struct GrayImageView {
    const std::uint8_t *data;
    std::uint32_t width, height;
};

namespace detail {
inline void check(rsvio_status s, rsvio_tracker *t, const char *what) {
    if (s != RSVIO_OK)
        throw Error(s, std::string(what) + ": " + rsvio_status_string(s) + (t ? std::string(" / ") + rsvio_tracker_last_error(t) : ""));
}
// status of a process_* call: a full track table is not an error -- the frame was processed and is delivered, only
// new points were dropped (the reference's maps are unbounded, its process_frame cannot fail); one warning on
// stderr per handle, the count stays queryable through rsvio_tracker_overflow_frames
inline void check_frame(rsvio_status s, rsvio_tracker *t, const char *what) {
    if (s == RSVIO_ERR_CAPACITY) {
        std::uint64_t n = 0;
        if (rsvio_tracker_overflow_frames(t, &n) == RSVIO_OK && n == 1)
            std::fprintf(stderr, "rsvio: %s: %s\n", what, rsvio_tracker_last_error(t));
        return;
    }
    check(s, t, what);
}

class Handle {
  public:
    Handle() = default;
    Handle(const Handle &) = delete;
    Handle &operator=(const Handle &) = delete;
    Handle(Handle &&o) noexcept : t_(o.t_), w_(o.w_), h_(o.h_) { o.t_ = nullptr; }
    ~Handle() {
        if (t_) rsvio_tracker_destroy(t_);
    }
    // the reference learns the image size from the first frame; the handle is created lazily the same way
    // returns true when a handle was (re)created
    bool ensure(const rsvio_config &base, std::uint32_t w, std::uint32_t h) {
        if (t_ && w == w_ && h == h_) return false;
        if (t_) rsvio_tracker_destroy(t_);
        t_ = nullptr;
        rsvio_config cfg = base;
        cfg.width = (std::int32_t)w;
        cfg.height = (std::int32_t)h;
        check(rsvio_tracker_create(&cfg, &t_), nullptr, "rsvio_tracker_create");
        w_ = w;
        h_ = h;
        return true;
    }
    rsvio_tracker *get() const { return t_; }
    std::unordered_map<std::size_t, std::pair<float, float>> tracks(int cam) const {
        std::unordered_map<std::size_t, std::pair<float, float>> out;
        if (!t_) return out;
        const rsvio_track_record *rec = nullptr;
        const std::int32_t *counts = nullptr;
        std::int32_t groups = 0, cap = 0;
        if (rsvio_tracker_result_view(t_, &rec, &counts, &groups, &cap) != RSVIO_OK) return out;  // no frame yet
        const int g = cam * (groups / 2);
        out.reserve((std::size_t)counts[g]);
        for (std::int32_t k = 0; k < counts[g]; ++k) {
            const rsvio_track_record &r = rec[(std::size_t)g * cap + k];
            out.emplace((std::size_t)r.id, std::make_pair(r.x, r.y));
        }
        return out;
    }
    void remove(const std::vector<std::size_t> &ids) {
        if (!t_ || ids.empty()) return;
        std::vector<std::uint64_t> v(ids.begin(), ids.end());
        check(rsvio_tracker_remove_ids(t_, 0, v.data(), v.size()), t_, "rsvio_tracker_remove_ids");
    }

  private:
    rsvio_tracker *t_ = nullptr;
    std::uint32_t w_ = 0, h_ = 0;
};
}  // namespace detail

template <std::uint32_t LEVELS>
class StereoPatchTracker {
  public:
    // StereoPatchTracker::new(grid_size, optical_flow_max_iterations, optical_flow_convergence_threshold)
    StereoPatchTracker(std::uint32_t grid_size, std::uint32_t optical_flow_max_iterations,
                       double optical_flow_convergence_threshold, int device = 0) {
        rsvio_config_default(&cfg_);
        cfg_.levels = (std::int32_t)LEVELS;
        cfg_.grid_size = grid_size;
        cfg_.max_iterations = (std::int32_t)optical_flow_max_iterations;
        cfg_.convergence_threshold = (float)optical_flow_convergence_threshold;
        cfg_.mode = RSVIO_MODE_STEREO;
        cfg_.device = device;
    }

    // process_frame(&mut self, greyscale_image0, greyscale_image1, frame)
    template <typename Frame>
    void process_frame(const GrayImageView &image0, const GrayImageView &image1, Frame &frame) {
        process(image0, image1);
        const rsvio_track_record *rec = nullptr;
        const std::int32_t *counts = nullptr;
        std::int32_t groups = 0, cap = 0;
        detail::check(rsvio_tracker_result_view(h_.get(), &rec, &counts, &groups, &cap), h_.get(), "result_view");
        const float *uxy = nullptr;  // unprojected coordinates, when set_camera() was called
        detail::check(rsvio_tracker_undistorted_view(h_.get(), &uxy, nullptr, nullptr), h_.get(), "undistorted_view");
        for (int cam = 0; cam < 2; ++cam)
            for (std::int32_t k = 0; k < counts[cam]; ++k) {
                const std::size_t i = (std::size_t)cam * cap + k;
                const rsvio_track_record &r = rec[i];
                Feature f((std::size_t)r.id, r.x, r.y);
                if (uxy) {
                    f.undistorted_coord[0] = uxy[2 * i];
                    f.undistorted_coord[1] = uxy[2 * i + 1];
                }
                if (cam == 0) frame.add_left_feature(f); else frame.add_right_feature(f);
            }
    }
    // Optional: the camera model of cam 0 (left) / 1 (right) (src/datasets/mod.rs:93-163).  Features then carry
    // undistorted_coord and Frame::add_*_feature can skip unproject_one (src/estimator/frame.rs:118,131).
    // Call it after the first frame has created the handle, or before: it is re-applied on creation.
    void set_camera(int cam, const rsvio_camera_model &model) {
        cams_[cam] = model;
        have_cam_[cam] = true;
        if (h_.get()) detail::check(rsvio_tracker_set_camera(h_.get(), cam, &model), h_.get(), "rsvio_tracker_set_camera");
    }
    void process(const GrayImageView &image0, const GrayImageView &image1) {
        if (image0.width != image1.width || image0.height != image1.height) throw Error(RSVIO_ERR_INVALID_ARG, "stereo images differ in size");
        const bool created = h_.ensure(cfg_, image0.width, image0.height);
        if (created)
            for (int cam = 0; cam < 2; ++cam)
                if (have_cam_[cam])
                    detail::check(rsvio_tracker_set_camera(h_.get(), cam, &cams_[cam]), h_.get(), "rsvio_tracker_set_camera");
        detail::check_frame(rsvio_tracker_process_stereo(h_.get(), image0.data, image1.data, image0.width), h_.get(),
                            "rsvio_tracker_process_stereo");
    }
    // the same for frames as the decoder delivers them (RSVIO_PIXEL_LUMA16 / RSVIO_PIXEL_RGB8): the
    // DynamicImage::to_luma8() of the reference's dataset players (src/datasets/euroc_player.rs:231) runs on the device
    void process_decoded(const void *image0, const void *image1, std::uint32_t width, std::uint32_t height,
                         std::size_t stride_bytes, std::int32_t pixel_format) {
        const bool created = h_.ensure(cfg_, width, height);
        if (created)
            for (int cam = 0; cam < 2; ++cam)
                if (have_cam_[cam])
                    detail::check(rsvio_tracker_set_camera(h_.get(), cam, &cams_[cam]), h_.get(), "rsvio_tracker_set_camera");
        detail::check_frame(rsvio_tracker_process_stereo_fmt(h_.get(), image0, image1, stride_bytes, pixel_format), h_.get(),
                            "rsvio_tracker_process_stereo_fmt");
    }
    // get_track_points(&self) -> [HashMap<usize, (f32, f32)>; 2]
    std::pair<std::unordered_map<std::size_t, std::pair<float, float>>, std::unordered_map<std::size_t, std::pair<float, float>>>
    get_track_points() const {
        return {h_.tracks(0), h_.tracks(1)};
    }
    // remove_id(&mut self, ids)
    void remove_id(const std::vector<std::size_t> &ids) { h_.remove(ids); }

  private:
    rsvio_config cfg_{};
    detail::Handle h_;
    rsvio_camera_model cams_[2]{};
    bool have_cam_[2] = {false, false};
};

template <std::uint32_t LEVELS>
class PatchTracker {
  public:
    explicit PatchTracker(int device = 0) {  // Default: 30 iterations, 0.005, grid 30 (feature_tracker.rs:51-52,65)
        rsvio_config_default(&cfg_);
        cfg_.levels = (std::int32_t)LEVELS;
        cfg_.mode = RSVIO_MODE_MONO;
        cfg_.device = device;
    }
    void process_frame(const GrayImageView &image) {
        h_.ensure(cfg_, image.width, image.height);
        detail::check_frame(rsvio_tracker_process_mono(h_.get(), image.data, image.width), h_.get(), "rsvio_tracker_process_mono");
    }
    std::unordered_map<std::size_t, std::pair<float, float>> get_track_points() const { return h_.tracks(0); }
    void remove_id(const std::vector<std::size_t> &ids) { h_.remove(ids); }

  private:
    rsvio_config cfg_{};
    detail::Handle h_;
};

// ---- crate B (the reference's standalone `feature_tracker` crate, row f-2) ----------------------------------------
// FeatureTrackingConfig / Feature / FeatureTracker / ext::Frame of feature_tracker/src/feature_tracker.rs:24-69,
// 77-193 and ext.rs:3-15, over rsvio_b_tracker_*.  Images are dense float arrays in [0, 1] (FloatGrayImage).
namespace crate_b {

struct FeatureTrackingConfig {
    std::size_t nlevels = 5;
    double ratio = 2.0;
    bool preprocessing_blur = true;
    float preprocessing_blur_sigma = 2.0f;
    float detection_threshold = 2.5f;
    std::uint32_t detection_min_dist = 15;
    float detection_blur = 6.0f;
    std::size_t optical_flow_max_iter = 25;
    float optical_flow_lm_lambda = 0.1f;
};

struct Feature {
    std::size_t feature_id;
    float central_point[2];
};

struct Frame {
    std::int64_t frame_id;
    std::vector<Feature> features;
    explicit Frame(std::int64_t id) : frame_id(id) {}
};

struct FloatGrayImageView {
    const float *data;
    std::int32_t width, height;
};

class FeatureTracker {
  public:
    explicit FeatureTracker(const FeatureTrackingConfig &config, int device = 0, int matching_cost = RSVIO_B_COST_SSD)
        : config_(config), device_(device), cost_(matching_cost) {}
    FeatureTracker(const FeatureTracker &) = delete;
    FeatureTracker &operator=(const FeatureTracker &) = delete;
    ~FeatureTracker() { rsvio_b_tracker_destroy(h_); }

    // process_frame(&mut self, in_image, frame): the frame comes back with the tracked features (old order) followed
    // by the new ones (fresh ids) and is kept as the previous frame
    void process_frame(const FloatGrayImageView &image, Frame frame) {
        ensure(image.width, image.height);
        const rsvio_status st = rsvio_b_tracker_process_frame(h_, image.data, frame.frame_id);
        if (st != RSVIO_OK && st != RSVIO_ERR_CAPACITY)  // a full table drops new corners, the frame is still valid
            throw Error(st, std::string("rsvio_b_tracker_process_frame: ") + rsvio_status_string(st));
        std::int32_t n = 0, n_new = 0;
        rsvio_b_tracker_num_features(h_, &n, &n_new);
        std::vector<std::uint64_t> ids((std::size_t)n);
        std::vector<float> xy(2 * (std::size_t)n);
        std::size_t got = 0;
        rsvio_b_tracker_get_features(h_, ids.data(), xy.data(), (std::size_t)n, &got);
        for (std::int32_t i = 0; i < n; ++i) frame.features.push_back(Feature{(std::size_t)ids[i], {xy[2 * i], xy[2 * i + 1]}});
        previous_frame_.assign(1, std::move(frame));
    }
    const Frame *last_frame() const { return previous_frame_.empty() ? nullptr : &previous_frame_[0]; }

    // get_pyramid(): levels of the last frame, finest first; empty before the first frame
    std::vector<std::vector<float>> get_pyramid(std::vector<std::pair<int, int>> *dims = nullptr) {
        std::vector<std::vector<float>> out;
        if (previous_frame_.empty()) return out;
        std::int32_t levels = 0, ws[RSVIO_MAX_LEVELS], hs[RSVIO_MAX_LEVELS];
        rsvio_b_tracker_pyramid_dims(h_, &levels, ws, hs);
        for (int l = 0; l < levels; ++l) {
            out.emplace_back((std::size_t)ws[l] * hs[l]);
            const rsvio_status st = rsvio_b_tracker_get_pyramid_level(h_, l, out.back().data());
            if (st != RSVIO_OK) throw Error(st, std::string("rsvio_b_tracker_get_pyramid_level: ") + rsvio_status_string(st));
            if (dims) dims->emplace_back(ws[l], hs[l]);
        }
        return out;
    }

  private:
    void ensure(std::int32_t w, std::int32_t h) {  // the size is learnt from the first frame
        if (h_ && w == w_ && h == hh_) return;
        rsvio_b_tracker_destroy(h_);
        h_ = nullptr;
        rsvio_b_config c;
        rsvio_b_config_default(&c);
        c.width = w;
        c.height = h;
        c.nlevels = (std::int32_t)config_.nlevels;
        c.ratio = config_.ratio;
        c.preprocessing_blur = config_.preprocessing_blur ? 1 : 0;
        c.preprocessing_blur_sigma = config_.preprocessing_blur_sigma;
        c.detection_threshold = config_.detection_threshold;
        c.detection_min_dist = config_.detection_min_dist;
        c.detection_blur = config_.detection_blur;
        c.optical_flow_max_iter = (std::int32_t)config_.optical_flow_max_iter;
        c.optical_flow_lm_lambda = config_.optical_flow_lm_lambda;
        c.matching_cost = cost_;
        c.device = device_;
        const rsvio_status st = rsvio_b_tracker_create(&c, &h_);
        if (st != RSVIO_OK) throw Error(st, std::string("rsvio_b_tracker_create: ") + rsvio_status_string(st));
        w_ = w;
        hh_ = h;
    }
    FeatureTrackingConfig config_;
    int device_, cost_;
    rsvio_b_tracker *h_ = nullptr;
    std::int32_t w_ = 0, hh_ = 0;
    std::vector<Frame> previous_frame_;
};

}  // namespace crate_b

}  // namespace rsvio
