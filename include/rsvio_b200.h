/*
 * rsvio_b200.h -- C ABI of the B200-native stereo patch tracker (librsvio_b200.so).
 *
 * Drop-in boundary for RS-VIO's tracker A, `src/feature_tracker/` of the reference
 * (all file:line citations are relative to the reference repository root).  The Rust
 * side (rs-vio_b200/rust/feature_tracker) keeps `StereoPatchTracker<N>` / `PatchTracker<N>`
 * with the reference's signatures and forwards to the functions below; INTEGRATION.md shows
 * the binding.  Plain pointers and sizes only; no CUDA, torch or C++ types cross this line.
 *
 * Conventions
 *  - every function returns an rsvio_status (0 = ok, < 0 = error); nothing unwinds or aborts
 *    across the boundary; rsvio_tracker_last_error() gives the text of the last failure.
 *  - per-point tracking failures are NOT errors: they are reported as status codes per point
 *    (stateless entry points) or by the point disappearing from the track table (stateful),
 *    exactly as the reference drops them (feature_tracker.rs:259-287, 316-330, 362-391).
 *  - images are row-major 8-bit luma, `stride` bytes between rows (GrayImage: stride == width).
 *  - a transform is the 6 free entries of the reference's `na::Affine2<f32>`:
 *    T = [m11, m12, m13(tx), m21, m22, m23(ty)]; the third row is (0, 0, 1).
 *  - one handle == one CUDA device + one CUDA stream towards the caller (batched handles use further
 *    internal streams to overlap consecutive frames; rsvio_tracker_sync orders everything the caller
 *    can observe); a handle is not re-entrant, distinct
 *    handles may be driven from distinct threads (mirrors `&mut self`, feature_tracker.rs:116).
 *    Every call runs on the handle's device and restores the calling thread's current device.
 *  - there is no CPU fallback: without a usable CUDA device every entry point fails with
 *    RSVIO_ERR_CUDA.
 */
#ifndef RSVIO_B200_H
#define RSVIO_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RSVIO_ABI_VERSION 1
#define RSVIO_PATTERN_SIZE 52 /* patch.rs:7 PATTERN52_SIZE */
#define RSVIO_MAX_LEVELS 8

typedef int32_t rsvio_status;
enum {
    RSVIO_OK = 0,
    RSVIO_ERR_INVALID_ARG = -1,
    RSVIO_ERR_CUDA = -2,      /* CUDA runtime error or no device */
    RSVIO_ERR_CAPACITY = -3,  /* a track table or output array was too small */
    RSVIO_ERR_STATE = -4,     /* call order violated (e.g. sync without submit) */
    RSVIO_ERR_ALLOC = -5
};

/* per-point outcome of one track_points element (feature_tracker.rs:259-287) */
enum {
    RSVIO_TRACK_OK = 0,
    RSVIO_FAIL_TEMPLATE = 1,  /* Pattern52::new invalid             patch.rs:146,156-158 */
    RSVIO_FAIL_RESIDUAL = 2,  /* residual() == None                 patch.rs:212,227 */
    RSVIO_FAIL_NONFINITE = 3, /* non-finite increment               feature_tracker.rs:366 */
    RSVIO_FAIL_DIVERGED = 4,  /* |inc| > 1e6                        feature_tracker.rs:369 */
    RSVIO_FAIL_OOB = 5,       /* left the image after an update     feature_tracker.rs:380-388 */
    RSVIO_FAIL_FB = 6,        /* forward-backward distance^2 >= 0.4 feature_tracker.rs:277-281 */
    RSVIO_FAIL_BACKWARD = 16  /* added to 1..5 when the backward pass was the one that failed */
};

enum { RSVIO_MODE_STEREO = 0, RSVIO_MODE_MONO = 1 };

/* Replaces the constructor arguments of StereoPatchTracker::<N>::new(grid_size,
 * optical_flow_max_iterations, optical_flow_convergence_threshold) (feature_tracker.rs:103-114,
 * called from estimator.rs:85-89) plus the const generic N (estimator.rs:27). */
typedef struct rsvio_config {
    int32_t width;                 /* image width in pixels (fixed for the handle's lifetime) */
    int32_t height;
    int32_t levels;                /* N of StereoPatchTracker<N>; level i is (w >> i, h >> i) */
    uint32_t grid_size;            /* feature_detection.grid_size (config/euroc_vio.yaml:42) */
    int32_t max_iterations;        /* optical_flow_max_iterations */
    float convergence_threshold;   /* optical_flow_convergence_threshold; must be finite, <= 0 never converges early
                                      (as `inc.norm() < thr`, feature_tracker.rs:374) */
    int32_t mode;                  /* RSVIO_MODE_STEREO | RSVIO_MODE_MONO (PatchTracker<N>) */
    int32_t n_streams;             /* independent sequences batched on this handle; 1 == reference */
    int32_t device;                /* CUDA device ordinal */
    int32_t max_tracks;            /* capacity per stream and camera; 0 = 4 * cells + 64 */
    void *cuda_stream;             /* optional cudaStream_t to run on; NULL = handle creates its own */
} rsvio_config;

typedef struct rsvio_tracker rsvio_tracker;

/* ---- library ------------------------------------------------------------------------- */
int32_t rsvio_abi_version(void);
const char *rsvio_status_string(rsvio_status s);
/* number of visible CUDA devices (0 and RSVIO_ERR_CUDA text in *err if none) */
rsvio_status rsvio_device_count(int32_t *n);
/* fills the defaults of PatchTracker (30 iterations, 0.005, grid 30: feature_tracker.rs:51-52,65) */
void rsvio_config_default(rsvio_config *cfg);

/* pinned host memory for image hand-off without a staging copy (ingest row f-3;
 * replaces the `to_vec()` copies at estimator.rs:131,143) */
rsvio_status rsvio_host_alloc(void **ptr, size_t bytes);
/* the same with RSVIO_HOST_WRITE_COMBINED: pages the CPU only ever writes (decoded frames on their way to the
 * device) -- the DMA engine reads them without snooping the CPU caches; CPU reads of such memory are very slow */
#define RSVIO_HOST_WRITE_COMBINED 1u
rsvio_status rsvio_host_alloc_flags(void **ptr, size_t bytes, uint32_t flags);
rsvio_status rsvio_host_free(void *ptr);

/* ---- stateful tracker: StereoPatchTracker<N> / PatchTracker<N> ------------------------- */
/* StereoPatchTracker::new (feature_tracker.rs:103) / PatchTracker::default (:35-41) */
rsvio_status rsvio_tracker_create(const rsvio_config *cfg, rsvio_tracker **out);
void rsvio_tracker_destroy(rsvio_tracker *t);
const char *rsvio_tracker_last_error(const rsvio_tracker *t);
/* drops all tracks and the previous pyramids; ids restart at 0 */
rsvio_status rsvio_tracker_reset(rsvio_tracker *t);

/* StereoPatchTracker::process_frame (feature_tracker.rs:116-187) for stream 0 of a
 * 1-stream handle: upload, 2 pyramids, temporal L->L and R->R, replenish, stereo L->R,
 * id merge.  Host pointers (pinned or pageable).  Synchronous. */
rsvio_status rsvio_tracker_process_stereo(rsvio_tracker *t, const uint8_t *left, const uint8_t *right,
                                          size_t stride);
/* Same, for a frame as the decoder delivers it (ingest row f-3): the dataset players of the reference call
 * `DynamicImage::to_luma8()` on the CPU before the tracker sees the frame (datasets/euroc_player.rs:227-231,
 * tum_vi_player.rs:227-231); here the raw buffer is uploaded and converted on the device.
 *   RSVIO_PIXEL_LUMA8  1 byte per pixel  (identical to rsvio_tracker_process_stereo)
 *   RSVIO_PIXEL_LUMA16 native-endian u16 (TUM-VI PNGs): (c + 128) / 257
 *   RSVIO_PIXEL_RGB8   r, g, b bytes:     (2126 r + 7152 g + 722 b) / 10000
 * (`image` 0.25 arithmetic, not vendored in the reference: PARITY UNPINNED.)  stride in BYTES. */
enum { RSVIO_PIXEL_LUMA8 = 0, RSVIO_PIXEL_LUMA16 = 1, RSVIO_PIXEL_RGB8 = 2 };
rsvio_status rsvio_tracker_process_stereo_fmt(rsvio_tracker *t, const void *left, const void *right,
                                              size_t stride_bytes, int32_t pixel_format);
/* PatchTracker::process_frame (feature_tracker.rs:43-77) */
rsvio_status rsvio_tracker_process_mono(rsvio_tracker *t, const uint8_t *image, size_t stride);

/* the same for all n_streams sequences of the handle in one batched pass: left[s], right[s]
 * are the images of stream s (right may be NULL in mono mode) */
rsvio_status rsvio_tracker_process_batch(rsvio_tracker *t, const uint8_t *const *left,
                                         const uint8_t *const *right, size_t stride);
/* asynchronous form: submit enqueues upload + kernels and returns; sync waits for the oldest
 * submitted frame.  At most 2 frames may be in flight (double-buffered inputs). */
rsvio_status rsvio_tracker_submit_batch(rsvio_tracker *t, const uint8_t *const *left,
                                        const uint8_t *const *right, size_t stride);
rsvio_status rsvio_tracker_sync(rsvio_tracker *t);
/* inputs already resident in HBM: n_streams images back to back, `pitch` bytes apart,
 * `stride` bytes between rows.  Asynchronous; rsvio_tracker_sync waits.  The images must be complete
 * in device memory when the call is made and stay untouched until the frame has been synced: on
 * batched handles the pyramid is built on an internal stream, beside the tail of the previous frame. */
rsvio_status rsvio_tracker_submit_device(rsvio_tracker *t, const void *d_left, const void *d_right,
                                         size_t stride, size_t pitch);

/* get_track_points (feature_tracker.rs:188-200): ids ascending, xy = n x (x, y) */
rsvio_status rsvio_tracker_num_tracks(rsvio_tracker *t, int32_t stream, int32_t cam, int32_t *n);
rsvio_status rsvio_tracker_get_tracks(rsvio_tracker *t, int32_t stream, int32_t cam, uint64_t *ids,
                                      float *xy, size_t cap, size_t *n);
/* zero-copy view of the last synced frame's result: the merge kernel writes these records
 * straight into host-mapped pinned memory, so populating `Frame` (feature_tracker.rs:177-186)
 * needs no further copy.  records[(cam * n_streams + stream) * capacity + k], k < counts[group].
 * Valid until the second-next submit; readable while at most one frame is in flight. */
typedef struct rsvio_track_record {
    uint64_t id;
    float x, y;
} rsvio_track_record;
rsvio_status rsvio_tracker_result_view(rsvio_tracker *t, const rsvio_track_record **records,
                                       const int32_t **counts, int32_t *groups, int32_t *capacity);
/* the full Affine2 of every track (tracked_points_map_cam{0,1}, feature_tracker.rs:93,95) */
rsvio_status rsvio_tracker_get_transforms(rsvio_tracker *t, int32_t stream, int32_t cam, uint64_t *ids,
                                          float *T6, size_t cap, size_t *n);
/* remove_id (feature_tracker.rs:201-206): removes from both cameras.  RSVIO_ERR_STATE while a submitted
 * frame has not been synced. */
rsvio_status rsvio_tracker_remove_ids(rsvio_tracker *t, int32_t stream, const uint64_t *ids, size_t n);

/* ---- next step of every feature: unprojection -------------------------------
 * Frame::add_left_feature / add_right_feature (estimator/frame.rs:107-134) fill
 * Feature::undistorted_coord with camera.unproject_one(pixel)[0..2]; the estimator consumes it as the
 * normalised image coordinate (x/z, y/z) (optimization/factors.rs:13-25).  With a camera model set,
 * the export kernel writes that coordinate for every live record next to the (id, x, y) records.
 * The arithmetic of the reference lives in the un-vendored crate camera-intrinsic-model 0.7.2:
 * PARITY UNPINNED (model equations restated from the literature, inverse solved to f64 convergence;
 * tools/reference_vectors dumps the crate's own answers for comparison).
 *   RSVIO_CAMERA_PINHOLE_RADTAN: params = fx fy cx cy k1 k2 p1 p2 k3   (datasets/mod.rs:115-127)
 *   RSVIO_CAMERA_EUCM:           params = fx fy cx cy alpha beta        (datasets/mod.rs:100-108)
 * Pixels outside the model's domain give NaN. */
enum { RSVIO_CAMERA_NONE = 0, RSVIO_CAMERA_PINHOLE_RADTAN = 1, RSVIO_CAMERA_EUCM = 2 };
typedef struct rsvio_camera_model {
    int32_t model;
    int32_t reserved;
    double params[9];
} rsvio_camera_model;
/* cam: 0 = left, 1 = right; model RSVIO_CAMERA_NONE switches the output off again.  RSVIO_ERR_STATE while a
 * submitted frame has not been synced.  Takes effect from the next frame. */
rsvio_status rsvio_tracker_set_camera(rsvio_tracker *t, int32_t cam, const rsvio_camera_model *model);
/* uxy[(group * capacity + k) * 2 + {0, 1}] for the records of rsvio_tracker_result_view (same groups, same
 * capacity, same validity); NULL in *uxy when no camera model has been set. */
rsvio_status rsvio_tracker_undistorted_view(rsvio_tracker *t, const float **uxy, int32_t *groups, int32_t *capacity);

/* counters since create/reset, summed over streams:
 * [0] tracks attempted (track_points elements)  [1] frames (per stream)
 * [2] LK iterations evaluated                   [3] Pattern52 templates built
 * [4] kernels launched by this handle           [5] points detected by grid replenishment */
rsvio_status rsvio_tracker_counters(rsvio_tracker *t, uint64_t out[6]);

/* RSVIO_ERR_CAPACITY from rsvio_tracker_sync / process_* is NOT fatal: the frame has been processed, its result
 * views are valid and the state has advanced; only the newly detected points that did not fit the track table
 * (rsvio_config.max_tracks) were dropped.  The reference's HashMaps are unbounded (feature_tracker.rs:93-96), so
 * its signatures cannot fail this way: callers (the Rust / C++ / Python mirrors do) should deliver the frame and
 * surface a warning.  *n = frames since create/reset that dropped points. */
rsvio_status rsvio_tracker_overflow_frames(rsvio_tracker *t, uint64_t *n);

/* per-stage device time (CUDA events on the handle's stream), accumulated since the last call
 * with reset != 0.  ms[0..7]: upload, pyramid, temporal LK, compaction+occupancy, detection,
 * stereo LK, merge, download.  launches[i] = kernel launches (or copies) in that stage. */
#define RSVIO_N_STAGES 8
rsvio_status rsvio_tracker_set_profiling(rsvio_tracker *t, int32_t enabled);
rsvio_status rsvio_tracker_stage_times(rsvio_tracker *t, float ms[RSVIO_N_STAGES],
                                       uint64_t launches[RSVIO_N_STAGES], int32_t reset);

/* ---- stateless entry points (parity tests, kernel benchmarks) -------------------------- */
/* bytes of a pyramid stored level after level, each level dense (stride == width) */
size_t rsvio_pyramid_bytes(int32_t width, int32_t height, int32_t levels);

/* bytes between two pyramids in the device layout (every level start aligned to 256 B) */
size_t rsvio_pyramid_device_bytes(int32_t width, int32_t height, int32_t levels);

/* build_image_pyramid (feature_tracker.rs:209-220) for n_images images that are `pitch`
 * bytes apart; out receives n_images pyramids back to back.  Host pointers. */
rsvio_status rsvio_build_pyramid_u8(int32_t device, const uint8_t *images, int32_t n_images,
                                    int32_t width, int32_t height, size_t stride, size_t pitch,
                                    int32_t levels, uint8_t *out);
/* same on device pointers; d_out holds n_images * rsvio_pyramid_device_bytes() (device layout).
 * `repeats` back-to-back launches timed with CUDA events on the launching stream;
 * kernel_ms = average milliseconds per launch (NULL to skip) */
rsvio_status rsvio_build_pyramid_u8_device(int32_t device, const void *d_images, int32_t n_images,
                                           int32_t width, int32_t height, size_t stride, size_t pitch,
                                           int32_t levels, void *d_out, int32_t repeats, float *kernel_ms);

/* track_points (feature_tracker.rs:252-291): n points from pyramid A to pyramid B
 * (rsvio_pyramid_bytes layout, host pointers), forward + backward + FB test.
 * out_T6: forward result (valid when status == 0); status: RSVIO_TRACK_* / RSVIO_FAIL_*;
 * iters (may be NULL): n x 2 LK iterations evaluated forward / backward. */
rsvio_status rsvio_track_points(int32_t device, const uint8_t *pyr_a, const uint8_t *pyr_b,
                                int32_t width, int32_t height, int32_t levels, int32_t n,
                                const float *in_T6, int32_t max_iterations, float convergence_threshold,
                                float *out_T6, int32_t *status, int32_t *iters);

/* detect_key_points with one point per cell (image_utilities.rs:108-175; add_points
 * feature_tracker.rs:222-251).  cur_xy: n_cur x 2 rounded positions of the tracked points.
 * out_xy / out_score hold up to cap corners in the reference's visiting order. */
rsvio_status rsvio_detect_keypoints(int32_t device, const uint8_t *image, int32_t width, int32_t height,
                                    size_t stride, uint32_t grid_size, const uint32_t *cur_xy, int32_t n_cur,
                                    uint32_t *out_xy, float *out_score, size_t cap, size_t *n);

/* se2_exp_matrix (image_utilities.rs:82-106) evaluated by the device code: n twists
 * [vx, vy, theta] -> n row-major 3x3 matrices */
rsvio_status rsvio_se2_exp(int32_t device, const float *twists, int32_t n, float *out9);

/* In-run FP32 microbenchmark (the denominator of the LK kernel's roofline): TFLOP/s of
 * dependent-chain FFMA and of separately rounded FMUL + FADD pairs (2 flop per multiply-add) on `device`, at the
 * clocks of the moment.  Takes a few milliseconds. */
rsvio_status rsvio_microbench_fp32(int32_t device, float *ffma_tflops, float *fmul_fadd_tflops);

/* unproject_one for n pixels (host pointers): xy = n x (u, v) float, out_xy = n x (x/z, y/z) float */
rsvio_status rsvio_unproject_points(int32_t device, const rsvio_camera_model *model, const float *xy, int32_t n,
                                    float *out_xy);

/* DynamicImage::to_luma8 for one decoded frame (host pointers); out = height x width bytes, dense */
rsvio_status rsvio_convert_to_luma8(int32_t device, const void *src, int32_t width, int32_t height,
                                    size_t stride_bytes, int32_t pixel_format, uint8_t *out);

/* ---- row f-2: the tracking core of the reference's second tracker (`feature_tracker` crate) ----
 * track_points (feature_tracker/src/feature_tracker/feature_tracking.rs:16-59) for n points between two f32
 * pyramids (host pointers; level l is heights[l] x widths[l] floats, dense): bicubic sampler, SSD patch of 52
 * samples (radius 7 px), LM-damped Gauss-Newton on SE(2) with twist [theta, vx, vy], forward-backward test
 * (< 2 px).  out_T4[i] = (cos, sin, tx, ty) of the forward isometry; status[i]: 0 tracked, 1 forward pass left
 * the image, 2 backward pass left the image, 3 forward-backward distance >= 2 px.  SSD cost. */
rsvio_status rsvio_b_track_points(int32_t device, const float *const *pyr0, const float *const *pyr1,
                                  const int32_t *widths, const int32_t *heights, int32_t levels,
                                  const float *points_xy, int32_t n, int32_t max_iteration, float lm_lambda,
                                  float *out_T4, int32_t *status);
/* crate B's build_image_pyramid without the preprocessing blur (feature_tracker/src/image_operations.rs:47-78):
 * level 0 = image (h x w floats in [0, 1]), level l = Triangle resize of level l-1 to
 * round(w / ratio^l) x round(h / ratio^l).  rsvio_b_pyramid_dims fills the level sizes; out_levels[l] must hold
 * heights[l] * widths[l] floats (host pointers). */
rsvio_status rsvio_b_pyramid_dims(int32_t width, int32_t height, int32_t levels, double ratio, int32_t *widths,
                                  int32_t *heights);
rsvio_status rsvio_b_build_pyramid(int32_t device, const float *image, int32_t width, int32_t height, int32_t levels,
                                   double ratio, float *const *out_levels);

/* ---- row f-2: the rest of crate B -- preprocessing blur, LSSD cost, Shi-Tomasi detection, the stateful tracker ----
 * matching cost of Patch52 (feature_tracker/src/patch.rs:4-8); FeatureTracker::process_frame hard-codes SSD
 * (feature_tracker.rs:125), LSSD is reachable through the public `patch` module */
#define RSVIO_B_COST_SSD 0
#define RSVIO_B_COST_LSSD 1
/* rsvio_b_track_points with the matching cost as a parameter */
rsvio_status rsvio_b_track_points_cost(int32_t device, const float *const *pyr0, const float *const *pyr1,
                                       const int32_t *widths, const int32_t *heights, int32_t levels,
                                       const float *points_xy, int32_t n, int32_t max_iteration, float lm_lambda,
                                       int32_t matching_cost, float *out_T4, int32_t *status);
/* `image::imageops::blur(image, sigma)` as build_image_pyramid applies it (image_operations.rs:58-62); host
 * pointers, height x width floats in [0, 1] */
rsvio_status rsvio_b_blur(int32_t device, const float *image, int32_t width, int32_t height, float sigma, float *out);
/* shi_tomasi_score (feature_tracker/src/feature_tracker/feature_detection.rs:83-165); host pointers */
rsvio_status rsvio_b_shi_tomasi_score(int32_t device, const float *image, int32_t width, int32_t height,
                                      float detection_blur, float *out_score);
/* add_points (feature_detection.rs:48-80) on a finest-level image: block non-maximum suppression (radius 1) of the
 * score map, the n_tracked tracked features injected with score max + 1, `local_maxima(min_dist)`, border filter.
 * out_xy = (x, y) u32 pairs of the new corners in (y, x) order; *n = their number (at most cap are written). */
rsvio_status rsvio_b_add_points(int32_t device, const float *image, int32_t width, int32_t height,
                                const float *tracked_xy, int32_t n_tracked, float threshold, uint32_t min_dist,
                                float detection_blur, uint32_t *out_xy, size_t cap, size_t *n);

/* FeatureTrackingConfig (feature_tracker/src/feature_tracker.rs:24-38) + the image size, device and matching cost */
typedef struct rsvio_b_config {
    int32_t width, height;              /* fixed for the handle's lifetime */
    int32_t nlevels;                    /* number of pyramid levels */
    double ratio;                       /* ratio of size between levels */
    int32_t preprocessing_blur;         /* bool */
    float preprocessing_blur_sigma;
    float detection_threshold;
    uint32_t detection_min_dist;
    float detection_blur;
    int32_t optical_flow_max_iter;
    float optical_flow_lm_lambda;
    int32_t matching_cost;              /* RSVIO_B_COST_SSD (what process_frame uses) | RSVIO_B_COST_LSSD */
    int32_t device;
    int32_t max_features;               /* capacity of the feature table; 0 = derived from min_dist */
} rsvio_b_config;
/* the values of feature_tracker/config/config.yaml (5 levels, ratio 2, blur 2.0, threshold 2.5, min_dist 15,
 * detection blur 6.0, 25 iterations, lambda 0.1), SSD */
void rsvio_b_config_default(rsvio_b_config *cfg);

typedef struct rsvio_b_tracker rsvio_b_tracker;
/* FeatureTracker::new (feature_tracker.rs:59-69; the viewer is dropped) */
rsvio_status rsvio_b_tracker_create(const rsvio_b_config *cfg, rsvio_b_tracker **out);
void rsvio_b_tracker_destroy(rsvio_b_tracker *t);
/* FeatureTracker::process_frame (feature_tracker.rs:77-185): image = height x width floats in [0, 1] (a
 * FloatGrayImage, dense).  Synchronous like the reference: returns when the frame's features are on the host. */
rsvio_status rsvio_b_tracker_process_frame(rsvio_b_tracker *t, const float *image, int64_t frame_id);
/* the frame handed back by process_frame (`previous_frame_pyramid.0.features`): tracked features first, in their old
 * order, then the new ones in (y, x) order with fresh ids; *n_new = how many of them are new */
rsvio_status rsvio_b_tracker_num_features(const rsvio_b_tracker *t, int32_t *n, int32_t *n_new);
rsvio_status rsvio_b_tracker_get_features(const rsvio_b_tracker *t, uint64_t *ids, float *xy, size_t cap, size_t *n);
/* get_pyramid (feature_tracker.rs:190-193): level sizes and one level of the last frame's pyramid (host pointer,
 * heights[l] x widths[l] floats); RSVIO_ERR_INVALID_ARG before the first frame */
rsvio_status rsvio_b_tracker_pyramid_dims(const rsvio_b_tracker *t, int32_t *levels, int32_t *widths, int32_t *heights);
rsvio_status rsvio_b_tracker_get_pyramid_level(rsvio_b_tracker *t, int32_t level, float *out);
/* kernel time of the last process_frame (CUDA events on the handle's stream, upload and download excluded) */
rsvio_status rsvio_b_tracker_last_frame_ms(const rsvio_b_tracker *t, float *ms);

#ifdef __cplusplus
}
#endif
#endif /* RSVIO_B200_H */
