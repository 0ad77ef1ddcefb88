// apib.cu -- C ABI of row f-2: the reference's second tracker (the `feature_tracker` crate, "crate B") as a stateful
// handle, plus the stateless entry points the parity tests drive.
//
// Replaces (paths relative to the reference repository):
//   FeatureTracker::new / process_frame / get_pyramid   feature_tracker/src/feature_tracker.rs:59-69, 77-185, 190-193
//   build_image_pyramid (blur + chained resize)          feature_tracker/src/image_operations.rs:47-78
//   feature_tracking::track_points                      feature_tracker/src/feature_tracker/feature_tracking.rs:16-59
//   feature_detection::add_points                       feature_tracker/src/feature_tracker/feature_detection.rs:48-80
// One frame = upload, blur, L - 1 resizes, one tracking launch (forward + backward, all levels), one compaction,
// the detection launches, one download; counts stay on the device in between (only the number of features of the
// previous frame, which the host returned to the caller anyway, sizes a grid).
#include <cmath>
#include <cstdio>
#include <cstring>

#include "common.cuh"

namespace rsvio {
namespace {

// ---- host-side f32 arithmetic of the un-vendored crates (restated; see oracle/crate_b.py) ----
// gaussian(x, r) = ((2 pi).sqrt() * r).recip() * (-x.powi(2) / (2.0 * r.powi(2))).exp(), support 2 sigma
bool make_gauss_tab(float sigma, GaussTab *G) {
    if (!(sigma > 0.0f)) sigma = 1.0f;  // imageops::blur: sigma <= 0 becomes 1
    const float support = 2.0f * sigma;
    const int K = (int)std::ceil(support + 0.5f) + 1;  // |i - o| <= ceil(o + 0.5 + support) - 1 - o
    if (K > 64) return false;
    const float norm = 1.0f / (std::sqrt(2.0f * 3.14159265358979323846f) * sigma);
    const float den = 2.0f * (sigma * sigma);
    for (int d = -K; d <= K; ++d) {
        const float x = (float)d;
        G->g[d + K] = norm * std::exp(-(x * x) / den);
    }
    G->K = K;
    G->support = support;
    return true;
}

// imageops/fast_blur.rs boxes_for_gauss(sigma, 3) -> box radii (size - 1) / 2
void box_radii(float sigma, int r[3]) {
    const float n = 3.0f;
    const float w_ideal = std::sqrt(12.0f * (sigma * sigma) / n + 1.0f);
    float w_l = std::floor(w_ideal);
    if (std::fmod(w_l, 2.0f) == 0.0f) w_l -= 1.0f;
    const float w_u = w_l + 2.0f;
    const float m_ideal = 0.25f * n * (w_l + 3.0f) - 3.0f * (sigma * sigma) * (1.0f / (w_l + 1.0f));
    long m = std::lround(m_ideal);  // f32::round: half away from zero
    if (m < 0) m = 0;
    for (int i = 0; i < 3; ++i) {
        const int size = (int)(i < m ? w_l : w_u);
        r[i] = size > 0 ? (size - 1) / 2 : 0;
    }
}

struct DevScope {
    int prev = -1, cur = -1;
    cudaError_t enter(int dev) {
        if (cudaGetDevice(&prev) != cudaSuccess) {
            cudaGetLastError();
            prev = -1;
        }
        cur = dev;
        return prev == dev ? cudaSuccess : cudaSetDevice(dev);
    }
    ~DevScope() {
        if (prev >= 0 && prev != cur) cudaSetDevice(prev);
    }
};

#define CUB(expr)                                                                                          \
    do {                                                                                                   \
        cudaError_t e__ = (expr);                                                                          \
        if (e__ != cudaSuccess) {                                                                          \
            fprintf(stderr, "rsvio: CUDA error %d (%s) at apib.cu:%d: %s\n", (int)e__, cudaGetErrorString(e__), \
                    __LINE__, #expr);                                                                      \
            cudaGetLastError();                                                                            \
            rs = RSVIO_ERR_CUDA;                                                                           \
            goto done;                                                                                     \
        }                                                                                                  \
    } while (0)

// detection scratch shared by the handle and the stateless calls
struct DetScratch {
    float *planes_a = nullptr, *planes_b = nullptr, *score = nullptr, *cand = nullptr;
    uint8_t *fmap = nullptr;
    uint32_t *list = nullptr, *surv = nullptr;
    unsigned *cnt = nullptr;
    cudaError_t alloc(int w, int h) {
        const size_t n = (size_t)w * h, half = (size_t)((w + 1) / 2) * ((h + 1) / 2);
        cudaError_t e;
        if ((e = cudaMalloc((void **)&planes_a, 3 * n * sizeof(float))) != cudaSuccess) return e;
        if ((e = cudaMalloc((void **)&planes_b, 3 * n * sizeof(float))) != cudaSuccess) return e;
        if ((e = cudaMalloc((void **)&score, n * sizeof(float))) != cudaSuccess) return e;
        if ((e = cudaMalloc((void **)&cand, n * sizeof(float))) != cudaSuccess) return e;
        if ((e = cudaMalloc((void **)&fmap, n)) != cudaSuccess) return e;
        if ((e = cudaMalloc((void **)&list, half * sizeof(uint32_t))) != cudaSuccess) return e;
        if ((e = cudaMalloc((void **)&surv, half * sizeof(uint32_t))) != cudaSuccess) return e;
        if ((e = cudaMalloc((void **)&cnt, DB_N * sizeof(unsigned))) != cudaSuccess) return e;
        return cudaMemset(cnt, 0, DB_N * sizeof(unsigned));
    }
    void release() {
        cudaFree(planes_a);
        cudaFree(planes_b);
        cudaFree(score);
        cudaFree(cand);
        cudaFree(fmap);
        cudaFree(list);
        cudaFree(surv);
        cudaFree(cnt);
    }
};

bool cfg_ok(const rsvio_b_config *c) {
    return c && c->width >= 16 && c->height >= 16 && c->width <= 65535 && c->height <= 65535 && c->nlevels >= 1 &&
           c->nlevels <= RSVIO_MAX_LEVELS && c->ratio > 1.0 && std::isfinite(c->preprocessing_blur_sigma) &&
           std::isfinite(c->detection_threshold) && c->detection_min_dist >= 1 &&
           c->detection_min_dist < 32768 && c->detection_blur > 0.0f && std::isfinite(c->detection_blur) &&
           c->optical_flow_max_iter >= 0 && std::isfinite(c->optical_flow_lm_lambda) &&
           (c->matching_cost == RSVIO_B_COST_SSD || c->matching_cost == RSVIO_B_COST_LSSD) && c->max_features >= 0;
}

}  // namespace
}  // namespace rsvio

using namespace rsvio;

struct rsvio_b_tracker {
    rsvio_b_config cfg;
    int w[RSVIO_MAX_LEVELS], h[RSVIO_MAX_LEVELS];
    cudaStream_t st = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    GaussTab gauss;
    int box_r[3];
    float *h_img = nullptr;                       // pinned staging copy of the caller's image
    float *d_in = nullptr, *d_tmp = nullptr;
    float *d_lv[2][RSVIO_MAX_LEVELS] = {};        // previous / current pyramid (ping-pong)
    int cur = 0;
    bool have_prev = false;
    int cap = 0;
    unsigned long long *d_ids[2] = {};
    float *d_xy[2] = {};
    int *d_n[2] = {};                             // feature counts of the two tables
    float *d_T4 = nullptr;
    int32_t *d_status = nullptr;
    unsigned long long *d_last_id = nullptr;
    int *d_n_new = nullptr, *d_err = nullptr;
    DetScratch det;
    // host copy of the last frame
    unsigned long long *h_ids = nullptr;
    float *h_xy = nullptr;
    int *h_counts = nullptr;                      // [0] n, [1] n_new, [2] overflow flag
    int n_feat = 0, n_new = 0;
    float last_ms = 0.0f;
};

extern "C" {

void rsvio_b_config_default(rsvio_b_config *cfg) {
    if (!cfg) return;
    memset(cfg, 0, sizeof(*cfg));
    cfg->nlevels = 5;
    cfg->ratio = 2.0;
    cfg->preprocessing_blur = 1;
    cfg->preprocessing_blur_sigma = 2.0f;
    cfg->detection_min_dist = 15;
    cfg->detection_threshold = 2.5f;
    cfg->detection_blur = 6.0f;
    cfg->optical_flow_max_iter = 25;
    cfg->optical_flow_lm_lambda = 0.1f;
    cfg->matching_cost = RSVIO_B_COST_SSD;
}

void rsvio_b_tracker_destroy(rsvio_b_tracker *t) {
    if (!t) return;
    DevScope ds;
    ds.enter(t->cfg.device);
    if (t->st) cudaStreamSynchronize(t->st);
    for (int k = 0; k < 2; ++k) {
        for (int l = 0; l < RSVIO_MAX_LEVELS; ++l) cudaFree(t->d_lv[k][l]);
        cudaFree(t->d_ids[k]);
        cudaFree(t->d_xy[k]);
        cudaFree(t->d_n[k]);
    }
    cudaFree(t->d_in);
    cudaFree(t->d_tmp);
    cudaFree(t->d_T4);
    cudaFree(t->d_status);
    cudaFree(t->d_last_id);
    cudaFree(t->d_n_new);
    cudaFree(t->d_err);
    t->det.release();
    if (t->h_img) cudaFreeHost(t->h_img);
    if (t->h_ids) cudaFreeHost(t->h_ids);
    if (t->h_xy) cudaFreeHost(t->h_xy);
    if (t->h_counts) cudaFreeHost(t->h_counts);
    if (t->ev0) cudaEventDestroy(t->ev0);
    if (t->ev1) cudaEventDestroy(t->ev1);
    if (t->st) cudaStreamDestroy(t->st);
    cudaGetLastError();
    delete t;
}

rsvio_status rsvio_b_tracker_create(const rsvio_b_config *cfg, rsvio_b_tracker **out) {
    if (!out || !cfg_ok(cfg)) return RSVIO_ERR_INVALID_ARG;
    *out = nullptr;
    rsvio_b_tracker *t = new rsvio_b_tracker();
    t->cfg = *cfg;
    rsvio_status rs = rsvio_b_pyramid_dims(cfg->width, cfg->height, cfg->nlevels, cfg->ratio, t->w, t->h);
    if (rs != RSVIO_OK || !make_gauss_tab(cfg->preprocessing_blur_sigma, &t->gauss)) {
        delete t;
        return RSVIO_ERR_INVALID_ARG;
    }
    box_radii(cfg->detection_blur, t->box_r);
    {
        // local_maxima keeps new corners more than min_dist apart (Chebyshev), tracked features may drift closer:
        // four times the packing bound, as tracker A sizes its tables
        const long cells = (long)(cfg->width / (int)cfg->detection_min_dist + 1) * (cfg->height / (int)cfg->detection_min_dist + 1);
        long cap = cfg->max_features > 0 ? cfg->max_features : 4 * cells + 64;
        if (cap > (1l << 22)) cap = 1l << 22;
        t->cap = (int)cap;
    }
    DevScope ds;
    const size_t n0 = (size_t)cfg->width * cfg->height;
    CUB(ds.enter(cfg->device));
    CUB(cudaStreamCreateWithFlags(&t->st, cudaStreamNonBlocking));
    CUB(cudaEventCreate(&t->ev0));
    CUB(cudaEventCreate(&t->ev1));
    CUB(cudaHostAlloc((void **)&t->h_img, n0 * sizeof(float), cudaHostAllocDefault));
    CUB(cudaMalloc((void **)&t->d_in, n0 * sizeof(float)));
    CUB(cudaMalloc((void **)&t->d_tmp, n0 * sizeof(float)));
    for (int k = 0; k < 2; ++k) {
        for (int l = 0; l < cfg->nlevels; ++l)
            CUB(cudaMalloc((void **)&t->d_lv[k][l], sizeof(float) * (size_t)t->w[l] * t->h[l]));
        CUB(cudaMalloc((void **)&t->d_ids[k], sizeof(unsigned long long) * t->cap));
        CUB(cudaMalloc((void **)&t->d_xy[k], sizeof(float) * 2 * t->cap));
        CUB(cudaMalloc((void **)&t->d_n[k], sizeof(int)));
        CUB(cudaMemset(t->d_n[k], 0, sizeof(int)));
    }
    CUB(cudaMalloc((void **)&t->d_T4, sizeof(float) * 4 * t->cap));
    CUB(cudaMalloc((void **)&t->d_status, sizeof(int32_t) * t->cap));
    CUB(cudaMalloc((void **)&t->d_last_id, sizeof(unsigned long long)));
    CUB(cudaMemset(t->d_last_id, 0, sizeof(unsigned long long)));
    CUB(cudaMalloc((void **)&t->d_n_new, sizeof(int)));
    CUB(cudaMalloc((void **)&t->d_err, sizeof(int)));
    CUB(cudaMemset(t->d_err, 0, sizeof(int)));
    CUB(t->det.alloc(cfg->width, cfg->height));
    CUB(cudaHostAlloc((void **)&t->h_ids, sizeof(unsigned long long) * t->cap, cudaHostAllocDefault));
    CUB(cudaHostAlloc((void **)&t->h_xy, sizeof(float) * 2 * t->cap, cudaHostAllocDefault));
    CUB(cudaHostAlloc((void **)&t->h_counts, sizeof(int) * 4, cudaHostAllocDefault));
    *out = t;
    return RSVIO_OK;
done:
    rsvio_b_tracker_destroy(t);
    return rs;
}

rsvio_status rsvio_b_tracker_process_frame(rsvio_b_tracker *t, const float *image, int64_t frame_id) {
    (void)frame_id;  // the reference only hands it to the viewer
    if (!t || !image) return RSVIO_ERR_INVALID_ARG;
    rsvio_status rs = RSVIO_OK;
    DevScope ds;
    const rsvio_b_config &c = t->cfg;
    const size_t n0 = (size_t)c.width * c.height;
    const int cur = t->cur ^ (t->have_prev ? 1 : 0), prev = cur ^ 1;
    const float *h_src = image;
    CUB(ds.enter(c.device));
    // a caller's pinned (cudaHostAlloc / rsvio_host_alloc / registered) image is DMA-read in place; pageable memory
    // goes through the handle's pinned staging copy (the call is synchronous: the buffer is free again on return)
    {
        cudaPointerAttributes pa;
        if (cudaPointerGetAttributes(&pa, image) != cudaSuccess || pa.type != cudaMemoryTypeHost) {
            cudaGetLastError();
            memcpy(t->h_img, image, n0 * sizeof(float));
            h_src = t->h_img;
        }
    }
    // build_image_pyramid: blur (optional), then level l = resize(level l - 1)
    if (c.preprocessing_blur) {
        CUB(cudaMemcpyAsync(t->d_in, h_src, n0 * sizeof(float), cudaMemcpyHostToDevice, t->st));
        CUB(cudaEventRecord(t->ev0, t->st));
        CUB(blurb_gaussian(t->d_in, c.width, c.height, t->gauss, t->d_tmp, t->d_lv[cur][0], t->st));
    } else {
        CUB(cudaMemcpyAsync(t->d_lv[cur][0], h_src, n0 * sizeof(float), cudaMemcpyHostToDevice, t->st));
        CUB(cudaEventRecord(t->ev0, t->st));
    }
    for (int l = 1; l < c.nlevels; ++l)
        CUB(resize_triangle_f32(t->d_lv[cur][l - 1], t->w[l - 1], t->h[l - 1], t->d_tmp, t->d_lv[cur][l], t->w[l],
                                t->h[l], t->st));
    // track_points previous -> current, tracked features in their old order
    if (t->have_prev && t->n_feat > 0) {
        CUB(trackb_points(t->d_lv[prev], t->d_lv[cur], t->w, t->h, c.nlevels, t->d_xy[prev], t->n_feat,
                          c.optical_flow_max_iter, c.optical_flow_lm_lambda, c.matching_cost, t->d_T4, t->d_status,
                          t->st));
        CUB(keepb_tracked(t->d_ids[prev], t->d_xy[prev], t->n_feat, t->d_T4, t->d_status, t->d_ids[cur], t->d_xy[cur],
                          t->d_n[cur], t->st));
    } else {
        CUB(cudaMemsetAsync(t->d_n[cur], 0, sizeof(int), t->st));
    }
    // add_points on the finest level
    CUB(scoreb_map(t->d_lv[cur][0], c.width, c.height, t->box_r, t->det.planes_a, t->det.planes_b, t->det.score, t->st));
    {
        DetectB d{};
        d.score = t->det.score;
        d.w = c.width;
        d.h = c.height;
        d.threshold = c.detection_threshold;
        d.min_dist = (int)c.detection_min_dist;
        d.cand = t->det.cand;
        d.fmap = t->det.fmap;
        d.list = t->det.list;
        d.surv = t->det.surv;
        d.cnt = t->det.cnt;
        d.ids = t->d_ids[cur];
        d.xy = t->d_xy[cur];
        d.n_feat = t->d_n[cur];
        d.n_tracked_host = t->have_prev ? t->n_feat : 0;
        d.cap = t->cap;
        d.last_id = t->d_last_id;
        d.n_new = t->d_n_new;
        d.err = t->d_err;
        CUB(detectb_new_points(d, t->st));
    }
    CUB(cudaEventRecord(t->ev1, t->st));
    // the frame's features -> host
    CUB(cudaMemcpyAsync(&t->h_counts[0], t->d_n[cur], sizeof(int), cudaMemcpyDeviceToHost, t->st));
    CUB(cudaMemcpyAsync(&t->h_counts[1], t->d_n_new, sizeof(int), cudaMemcpyDeviceToHost, t->st));
    CUB(cudaMemcpyAsync(&t->h_counts[2], t->d_err, sizeof(int), cudaMemcpyDeviceToHost, t->st));
    CUB(cudaStreamSynchronize(t->st));
    t->n_feat = t->h_counts[0];
    t->n_new = t->h_counts[1];
    if (t->n_feat > 0) {
        CUB(cudaMemcpyAsync(t->h_ids, t->d_ids[cur], sizeof(unsigned long long) * t->n_feat, cudaMemcpyDeviceToHost, t->st));
        CUB(cudaMemcpyAsync(t->h_xy, t->d_xy[cur], sizeof(float) * 2 * t->n_feat, cudaMemcpyDeviceToHost, t->st));
        CUB(cudaStreamSynchronize(t->st));
    }
    CUB(cudaEventElapsedTime(&t->last_ms, t->ev0, t->ev1));
    t->cur = cur;
    t->have_prev = true;
    if (t->h_counts[2]) {  // table overflow: the frame was processed, the excess new corners were dropped
        CUB(cudaMemsetAsync(t->d_err, 0, sizeof(int), t->st));
        return RSVIO_ERR_CAPACITY;
    }
done:
    return rs;
}

rsvio_status rsvio_b_tracker_num_features(const rsvio_b_tracker *t, int32_t *n, int32_t *n_new) {
    if (!t || !n) return RSVIO_ERR_INVALID_ARG;
    *n = t->n_feat;
    if (n_new) *n_new = t->n_new;
    return RSVIO_OK;
}

rsvio_status rsvio_b_tracker_get_features(const rsvio_b_tracker *t, uint64_t *ids, float *xy, size_t cap, size_t *n) {
    if (!t || !n || (cap > 0 && (!ids || !xy))) return RSVIO_ERR_INVALID_ARG;
    const size_t m = (size_t)t->n_feat < cap ? (size_t)t->n_feat : cap;
    for (size_t i = 0; i < m; ++i) ids[i] = (uint64_t)t->h_ids[i];
    if (m) memcpy(xy, t->h_xy, sizeof(float) * 2 * m);
    *n = (size_t)t->n_feat;
    return RSVIO_OK;
}

rsvio_status rsvio_b_tracker_pyramid_dims(const rsvio_b_tracker *t, int32_t *levels, int32_t *widths, int32_t *heights) {
    if (!t || !levels) return RSVIO_ERR_INVALID_ARG;
    *levels = t->cfg.nlevels;
    for (int l = 0; l < t->cfg.nlevels; ++l) {
        if (widths) widths[l] = t->w[l];
        if (heights) heights[l] = t->h[l];
    }
    return RSVIO_OK;
}

rsvio_status rsvio_b_tracker_get_pyramid_level(rsvio_b_tracker *t, int32_t level, float *out) {
    if (!t || !out || level < 0 || level >= t->cfg.nlevels || !t->have_prev) return RSVIO_ERR_INVALID_ARG;
    rsvio_status rs = RSVIO_OK;
    DevScope ds;
    CUB(ds.enter(t->cfg.device));
    CUB(cudaMemcpyAsync(out, t->d_lv[t->cur][level], sizeof(float) * (size_t)t->w[level] * t->h[level],
                        cudaMemcpyDeviceToHost, t->st));
    CUB(cudaStreamSynchronize(t->st));
done:
    return rs;
}

rsvio_status rsvio_b_tracker_last_frame_ms(const rsvio_b_tracker *t, float *ms) {
    if (!t || !ms) return RSVIO_ERR_INVALID_ARG;
    *ms = t->last_ms;
    return RSVIO_OK;
}

// ---- stateless pieces (parity tests) ---------------------------------------------------------------------------
rsvio_status rsvio_b_blur(int32_t device, const float *image, int32_t width, int32_t height, float sigma, float *out) {
    GaussTab G;
    if (!image || !out || width < 1 || height < 1 || !std::isfinite(sigma) || !make_gauss_tab(sigma, &G))
        return RSVIO_ERR_INVALID_ARG;
    rsvio_status rs = RSVIO_OK;
    DevScope ds;
    cudaStream_t st = nullptr;
    float *d = nullptr;
    const size_t n = (size_t)width * height;
    CUB(ds.enter(device));
    CUB(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    CUB(cudaMalloc((void **)&d, 3 * n * sizeof(float)));
    CUB(cudaMemcpyAsync(d, image, n * sizeof(float), cudaMemcpyHostToDevice, st));
    CUB(blurb_gaussian(d, width, height, G, d + n, d + 2 * n, st));
    CUB(cudaMemcpyAsync(out, d + 2 * n, n * sizeof(float), cudaMemcpyDeviceToHost, st));
    CUB(cudaStreamSynchronize(st));
done:
    cudaFree(d);
    if (st) cudaStreamDestroy(st);
    return rs;
}

rsvio_status rsvio_b_shi_tomasi_score(int32_t device, const float *image, int32_t width, int32_t height,
                                      float detection_blur, float *out_score) {
    if (!image || !out_score || width < 2 || height < 2 || !(detection_blur > 0.0f) || !std::isfinite(detection_blur))
        return RSVIO_ERR_INVALID_ARG;
    rsvio_status rs = RSVIO_OK;
    DevScope ds;
    cudaStream_t st = nullptr;
    float *d = nullptr;
    int r[3];
    box_radii(detection_blur, r);
    const size_t n = (size_t)width * height;
    CUB(ds.enter(device));
    CUB(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    CUB(cudaMalloc((void **)&d, 8 * n * sizeof(float)));
    CUB(cudaMemcpyAsync(d, image, n * sizeof(float), cudaMemcpyHostToDevice, st));
    CUB(scoreb_map(d, width, height, r, d + n, d + 4 * n, d + 7 * n, st));
    CUB(cudaMemcpyAsync(out_score, d + 7 * n, n * sizeof(float), cudaMemcpyDeviceToHost, st));
    CUB(cudaStreamSynchronize(st));
done:
    cudaFree(d);
    if (st) cudaStreamDestroy(st);
    return rs;
}

rsvio_status rsvio_b_add_points(int32_t device, const float *image, int32_t width, int32_t height,
                                const float *tracked_xy, int32_t n_tracked, float threshold, uint32_t min_dist,
                                float detection_blur, uint32_t *out_xy, size_t cap, size_t *n_out) {
    if (!image || !n_out || width < 2 || height < 2 || width > 65535 || height > 65535 || n_tracked < 0 ||
        (n_tracked > 0 && !tracked_xy) || min_dist < 1 || min_dist >= 32768 || !(detection_blur > 0.0f) ||
        !std::isfinite(detection_blur) || !std::isfinite(threshold) || (cap > 0 && !out_xy))
        return RSVIO_ERR_INVALID_ARG;
    rsvio_status rs = RSVIO_OK;
    DevScope ds;
    cudaStream_t st = nullptr;
    DetScratch det;
    float *d_img = nullptr, *d_xy = nullptr;
    unsigned long long *d_ids = nullptr, *d_last = nullptr;
    int *d_small = nullptr;  // [0] n_feat, [1] n_new, [2] err
    int r[3];
    int h_small[3] = {0, 0, 0};
    box_radii(detection_blur, r);
    const size_t n = (size_t)width * height;
    const int table = n_tracked + (int)(((size_t)(width + 1) / 2) * ((size_t)(height + 1) / 2));
    float *h_xy = nullptr;
    CUB(ds.enter(device));
    CUB(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    CUB(det.alloc(width, height));
    CUB(cudaMalloc((void **)&d_img, n * sizeof(float)));
    CUB(cudaMalloc((void **)&d_xy, sizeof(float) * 2 * (size_t)table));
    CUB(cudaMalloc((void **)&d_ids, sizeof(unsigned long long) * (size_t)table));
    CUB(cudaMalloc((void **)&d_last, sizeof(unsigned long long)));
    CUB(cudaMalloc((void **)&d_small, 3 * sizeof(int)));
    CUB(cudaMemsetAsync(d_last, 0, sizeof(unsigned long long), st));
    h_small[0] = n_tracked;
    CUB(cudaMemcpyAsync(d_small, h_small, 3 * sizeof(int), cudaMemcpyHostToDevice, st));
    CUB(cudaMemcpyAsync(d_img, image, n * sizeof(float), cudaMemcpyHostToDevice, st));
    if (n_tracked > 0) CUB(cudaMemcpyAsync(d_xy, tracked_xy, sizeof(float) * 2 * (size_t)n_tracked, cudaMemcpyHostToDevice, st));
    CUB(scoreb_map(d_img, width, height, r, det.planes_a, det.planes_b, det.score, st));
    {
        DetectB d{};
        d.score = det.score;
        d.w = width;
        d.h = height;
        d.threshold = threshold;
        d.min_dist = (int)min_dist;
        d.cand = det.cand;
        d.fmap = det.fmap;
        d.list = det.list;
        d.surv = det.surv;
        d.cnt = det.cnt;
        d.ids = d_ids;
        d.xy = d_xy;
        d.n_feat = d_small;
        d.n_tracked_host = n_tracked;
        d.cap = table;
        d.last_id = d_last;
        d.n_new = d_small + 1;
        d.err = d_small + 2;
        CUB(detectb_new_points(d, st));
    }
    CUB(cudaMemcpyAsync(h_small, d_small, 3 * sizeof(int), cudaMemcpyDeviceToHost, st));
    CUB(cudaStreamSynchronize(st));
    *n_out = (size_t)h_small[1];
    {
        const size_t m = (size_t)h_small[1] < cap ? (size_t)h_small[1] : cap;
        if (m) {
            h_xy = (float *)malloc(sizeof(float) * 2 * m);
            if (!h_xy) {
                rs = RSVIO_ERR_INVALID_ARG;
                goto done;
            }
            CUB(cudaMemcpyAsync(h_xy, d_xy + 2 * (size_t)n_tracked, sizeof(float) * 2 * m, cudaMemcpyDeviceToHost, st));
            CUB(cudaStreamSynchronize(st));
            for (size_t i = 0; i < 2 * m; ++i) out_xy[i] = (uint32_t)h_xy[i];
        }
    }
done:
    free(h_xy);
    det.release();
    cudaFree(d_img);
    cudaFree(d_xy);
    cudaFree(d_ids);
    cudaFree(d_last);
    cudaFree(d_small);
    if (st) cudaStreamDestroy(st);
    return rs;
}

}  // extern "C"
