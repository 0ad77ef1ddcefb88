// common.cuh -- shared host/device declarations of librsvio_b200 (sm_100a only).
//
// Data layout in HBM:
//   pyramid of one image: levels back to back, level i dense row-major (stride == w >> i),
//   each level start aligned to 256 B; `pyr_stride` bytes between the pyramids of two slots.
//   slot = cam * n_streams + stream.
//   track table of one (cam, stream) group: ids[cap] (u64) and T6[cap][6] (f32,
//   [m11,m12,tx,m21,m22,ty] of the reference's Affine2), count in counts[group].
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/rsvio_b200.h"
#include "camera.cuh"

namespace rsvio {

constexpr int kMaxLevels = RSVIO_MAX_LEVELS;
constexpr int kPattern = RSVIO_PATTERN_SIZE;
constexpr int kNumSMs = 148;  // B200 (fallback only: launches size their grids with sm_count())

// SM count of the calling thread's current device (grids are sized in multiples of it)
inline int sm_count() {
    static int cached[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) {
        cudaGetLastError();
        return kNumSMs;
    }
    if (cached[dev] <= 0) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) {
            cudaGetLastError();
            return kNumSMs;
        }
        cached[dev] = n;
    }
    return cached[dev];
}

struct LevelDesc {
    int w, h;
    uint32_t off;  // byte offset of the level inside one pyramid
};

struct PyrDesc {
    LevelDesc lv[kMaxLevels];
    int levels;
    uint32_t bytes;       // dense-with-alignment size of one pyramid
    uint32_t pyr_stride;  // bytes between consecutive slots (>= bytes, multiple of 256)
};

inline PyrDesc make_pyr_desc(int w, int h, int levels) {
    PyrDesc d{};
    d.levels = levels;
    uint32_t off = 0;
    for (int i = 0; i < levels; ++i) {
        d.lv[i].w = w >> i;
        d.lv[i].h = h >> i;
        d.lv[i].off = off;
        off += (uint32_t)(d.lv[i].w * d.lv[i].h);
        off = (off + 255u) & ~255u;
    }
    d.bytes = off;
    d.pyr_stride = off;
    return d;
}

// separable resampling taps of one (level, axis): image 0.25 `resize` weight construction
struct TapTable {
    const int *left;   // [n_out] first source index
    const int *cnt;    // [n_out] number of taps
    const float *w;    // [n_out][kmax] normalised weights
    int n_in, n_out, kmax;
};

struct PyrTaps {
    TapTable x[kMaxLevels], y[kMaxLevels];  // index 0 unused (level 0 is a copy)
};

// device counters (u64): see rsvio_tracker_counters
enum { CNT_TRACKS = 0, CNT_FRAMES, CNT_ITERS, CNT_TEMPLATES, CNT_LAUNCHES, CNT_DETECTED, CNT_N };

// ---- launch wrappers implemented in the .cu files -------------------------------------

// pyramid.cu
struct PyrPlan;  // owns the tap tables and tile lists of one (w, h, levels)
cudaError_t pyr_plan_create(int w, int h, int levels, cudaStream_t st, PyrPlan **out);
void pyr_plan_destroy(PyrPlan *p);
const PyrDesc &pyr_plan_desc(const PyrPlan *p);
// which of the plan's two tile counters the next fast-path launch draws from (a launch flips it): a captured launch
// sequence that is replayed must see the parity it was captured with, and the host copy must follow the replays
int pyr_plan_parity(const PyrPlan *p);
void pyr_plan_set_parity(PyrPlan *p, int parity);
// scratch for the deep dyadic levels (>= 4) of up to n_images images per pyr_build call
cudaError_t pyr_plan_reserve(PyrPlan *p, int n_images);
// n_images source images: image j at src[j < split ? 0 : 1] + (j < split ? j : j - split) * pitch
cudaError_t pyr_build(const PyrPlan *p, const uint8_t *src0, const uint8_t *src1, int split, int n_images,
                      size_t stride, size_t pitch, uint8_t *dst, cudaStream_t st, int *n_launches);

// lk.cu
struct LkBatch {
    const uint8_t *pyr_a;   // slot 0 of the source pyramids
    const uint8_t *pyr_b;   // slot 0 of the target pyramids
    const int *counts;      // [groups] points per group (device): the launch tracks rows [first[g], counts[g])
    const int *first;       // [groups] first row per group (device) or nullptr = 0
    const float *in_T6;     // [groups][cap][6]
    float *out_T6;          // [groups][cap][6]
    int32_t *status;        // [groups][cap]
    int32_t *iters;         // [groups][cap][2] or nullptr
    unsigned long long *counters;  // CNT_* or nullptr
    unsigned *work_counter;        // 2 device counters for dynamic point scheduling (zero at first use) or nullptr
    int work_parity;               // which of the two this launch uses; it clears the other for the next launch
    int groups, cap;
    int max_it;
    float thr;
    int blocks_per_sm;      // persistent CTAs per SM of a dynamically scheduled launch (0: as many as fit)
    uint32_t floor_bias;  // set by lk_track: the bit pattern of 2^23, passed as a launch parameter so that the
                          // compiler cannot split the folded address constant back into per-load immediates
};
cudaError_t lk_track(const PyrDesc &pd, const LkBatch &b, cudaStream_t st);
cudaError_t se2_exp_batch(const float *d_twists, int n, float *d_out9, cudaStream_t st);

// microbench.cu: measured FP32 multiply-add rates (TFLOP/s), fused and as separately rounded FMUL + FADD
cudaError_t fp32_rate(float *ffma_tflops, float *fmul_fadd_tflops, cudaStream_t st);

// detect.cu
struct GridDesc {
    int w, h;
    uint32_t g;
    int ncx, ncy;          // cells visited: w / g, h / g
    int occ_cols, occ_rows;  // w / g + 1, h / g + 1
    uint32_t x_start, y_start;
};
inline GridDesc make_grid_desc(int w, int h, uint32_t g) {
    GridDesc d{};
    d.w = w;
    d.h = h;
    d.g = g;
    d.ncx = (int)((uint32_t)w / g);
    d.ncy = (int)((uint32_t)h / g);
    d.occ_cols = d.ncx + 1;
    d.occ_rows = d.ncy + 1;
    d.x_start = ((uint32_t)w % g) / 2;
    d.y_start = ((uint32_t)h % g) / 2;
    return d;
}
// Detection work shared by the compaction kernel (producer of the list) and the detection kernel:
//   cell_pt [streams][cells]  packed (y << 16 | x) of the chosen corner or ~0; cell_score the same layout
//   list / list_count         empty cells of all streams (item = stream * cells + cell), count; next_count is the other
//                             counter of the pair (cleared for the next frame)
//   n_empty [streams]         listed cells per stream; done [streams] counts finished ones (zero between frames):
//                             the CTA that finishes a stream's last cell compacts that stream's new points
struct CellWork {
    uint32_t *cell_pt;
    float *cell_score;
    uint32_t *list;
    int *list_count, *next_count;
    int *n_empty, *done;
    int *new_count;  // [streams]
    int *arrive;     // [streams] chunks of the compaction launch that have marked their part of the occupancy grid
};
// occ: [streams][occ_rows*occ_cols] bytes.  img of stream s = img0 + s * pitch.  Stateless form: lists the empty cells
// itself (1 extra launch) before the detection launch.
cudaError_t detect_cells(const GridDesc &gd, const uint8_t *img0, size_t stride, size_t pitch, int streams,
                         const uint8_t *occ, const CellWork &cw, float *new_T6, float *new_score, int new_cap,
                         unsigned long long *counters, bool list_first, bool narrow, cudaStream_t st);
// marks cells holding one of n (x, y) rounded positions (stateless API); occ must be zeroed
cudaError_t mark_occupancy_xy(const GridDesc &gd, const uint32_t *d_xy, int n, uint8_t *occ, cudaStream_t st);

// tracks.cu
struct TableView {
    unsigned long long *ids;  // [groups][cap]
    float *T6;                // [groups][cap][6]
    int *counts;              // [groups]
};
// stable compaction of the surviving tracks of every group from `in` (+ lk results) into `out`;
// groups < streams (camera 0) also rebuild the occupancy grid of their stream.
// surv (optional): [groups] survivors per group, i.e. out.counts before the merge kernel appends the new points.
cudaError_t compact_tracks(const TableView &in, const float *lk_T6, const int32_t *lk_status, const TableView &out,
                           int groups, int cap, int streams, const GridDesc &gd, uint8_t *occ, int have_prev,
                           unsigned long long *counters, int *err_clear, const CellWork &cw, int *surv, cudaStream_t st);
// packed output record written to (mapped) host memory
struct OutRecord {
    unsigned long long id;
    float x, y;
};
// appends the new stereo pairs (status ok) to both cameras of `tab` and assigns their ids
cudaError_t merge_new_points(const TableView &tab, int streams, int cap, int stereo, const float *new_T6,
                             const float *st_T6, const int32_t *st_status, const int *new_count, int new_cap,
                             unsigned long long *last_id, int *err_flag, cudaStream_t st);
// packed (id, x, y) records and counts of every group -> out_rec / out_counts (mapped host memory); with a camera
// model (groups < streams: cam0, else cam1) also the unprojected coordinate of every record -> out_uxy
cudaError_t export_records(const TableView &tab, int groups, int cap, int streams, OutRecord *out_rec, int *out_counts,
                           const CameraModelDev &cam0, const CameraModelDev &cam1, float *out_uxy, const int *err_flag,
                           cudaStream_t st);  // err_flag (device, may be null) -> out_counts[groups]
cudaError_t unproject_points(const CameraModelDev &cm, const float *d_xy, int n, float *d_out, cudaStream_t st);
cudaError_t remove_ids(const TableView &in, const TableView &out, int groups, int cap, int streams, int stream,
                       const unsigned long long *d_ids, int n, cudaStream_t st);

}  // namespace rsvio

// ---- device helpers ---------------------------------------------------------------------
#ifdef __CUDACC__
// Bounds-check build (make VARIANT=check EXTRA=-DRSVIO_BOUNDS_CHECK): every unclamped gather, shared-memory tile index
// and scratch index of the hot kernels is asserted; a violation prints its source line and traps the kernel.  This is
// the repo's own stand-in for compute-sanitizer memcheck (closed on the GPU pool); tools/bounds_check.sh runs the
// parity tests against that build.  The default build compiles the checks away.
#ifdef RSVIO_BOUNDS_CHECK
#include <cstdio>
#define RSVIO_CHECK(cond)                                                                                       \
    do {                                                                                                        \
        if (!(cond)) {                                                                                          \
            printf("rsvio bounds check failed: %s at %s:%d (block %d,%d thread %d)\n", #cond, __FILE__, __LINE__, \
                   (int)blockIdx.x, (int)blockIdx.y, (int)threadIdx.x);                                         \
            __trap();                                                                                           \
        }                                                                                                       \
    } while (0)
#else
#define RSVIO_CHECK(cond) ((void)0)
#endif
namespace rsvio {

// Rust `f32 as u32`: saturating, NaN -> 0
__device__ __forceinline__ uint32_t sat_u32(float v) {
    return __float2uint_rz(v);  // cvt.rzi.u32.f32 saturates and maps NaN to 0
}

// inbound -- image_utilities.rs:75-80 (round half away from zero, then u32 compare)
__device__ __forceinline__ bool inbound(int w, int h, float x, float y, uint32_t r) {
    const uint32_t rx = sat_u32(roundf(x)), ry = sat_u32(roundf(y));
    return rx >= r && ry >= r && rx < (uint32_t)w - r && ry < (uint32_t)h - r;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Sums of several per-lane values over the warp with fewer shuffles than one butterfly per value:
// each xor step keeps half of the values and hands the other half to the partner lane, so 4 values
// need 2 + 1 + 3 shuffles (instead of 20) and end up on lanes 0, 8, 16, 24, from where they are
// broadcast.  Every lane returns the same bits.
__device__ __forceinline__ void warp_sum4(float &v0, float &v1, float &v2, float &v3, int lane) {
    const bool up16 = lane & 16, up8 = lane & 8;
    float a = up16 ? v2 : v0, b = up16 ? v3 : v1;
    a += __shfl_xor_sync(0xffffffffu, up16 ? v0 : v2, 16);
    b += __shfl_xor_sync(0xffffffffu, up16 ? v1 : v3, 16);
    float c = (up8 ? b : a) + __shfl_xor_sync(0xffffffffu, up8 ? a : b, 8);
    c += __shfl_xor_sync(0xffffffffu, c, 4);
    c += __shfl_xor_sync(0xffffffffu, c, 2);
    c += __shfl_xor_sync(0xffffffffu, c, 1);
    v0 = __shfl_sync(0xffffffffu, c, 0);
    v1 = __shfl_sync(0xffffffffu, c, 8);
    v2 = __shfl_sync(0xffffffffu, c, 16);
    v3 = __shfl_sync(0xffffffffu, c, 24);
}

// 8 values: 4 + 2 + 1 + 2 shuffles, results on lanes 0, 4, .., 28
__device__ __forceinline__ void warp_sum8(float (&v)[8], int lane) {
    const bool up16 = lane & 16, up8 = lane & 8, up4 = lane & 4;
    float a[4];
#pragma unroll
    for (int k = 0; k < 4; ++k)
        a[k] = (up16 ? v[4 + k] : v[k]) + __shfl_xor_sync(0xffffffffu, up16 ? v[k] : v[4 + k], 16);
    float b[2];
#pragma unroll
    for (int k = 0; k < 2; ++k)
        b[k] = (up8 ? a[2 + k] : a[k]) + __shfl_xor_sync(0xffffffffu, up8 ? a[k] : a[2 + k], 8);
    float c = (up4 ? b[1] : b[0]) + __shfl_xor_sync(0xffffffffu, up4 ? b[0] : b[1], 4);
    c += __shfl_xor_sync(0xffffffffu, c, 2);
    c += __shfl_xor_sync(0xffffffffu, c, 1);
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] = __shfl_sync(0xffffffffu, c, 4 * k);
}

// decoded frames (16-bit grey, RGB8) -> dense 8-bit luma (DynamicImage::to_luma8 of the dataset players)
int pixel_bytes(int fmt);
cudaError_t convert_to_luma8(const uint8_t *d_src, size_t stride_bytes, int w, int h, int fmt, uint8_t *d_dst,
                             cudaStream_t st);

// row f-2, first slice: crate B's tracking core (trackb.cu).  Device pointers; out_T4 = n x (cos, sin, tx, ty)
// cost: RSVIO_B_COST_SSD / RSVIO_B_COST_LSSD
cudaError_t trackb_points(const float *const *d_a, const float *const *d_b, const int *w, const int *h, int levels,
                          const float *d_pts, int n, int max_it, float lambda, int cost, float *d_out_T4,
                          int32_t *d_status, cudaStream_t st);

// one level of crate B's pyramid: Triangle resize of an f32 image (d_tmp: h_out x w_in floats of scratch)
cudaError_t resize_triangle_f32(const float *d_src, int w_in, int h_in, float *d_tmp, float *d_dst, int w_out, int h_out,
                                cudaStream_t st);


// ---- detectb.cu: crate B's blur, Shi-Tomasi detection and feature bookkeeping (device pointers) ----
struct GaussTab {     // un-normalised gaussian(d, sigma) for d = -K .. K, made on the host (libm expf); support = 2 sigma
    float g[2 * 64 + 1];
    int K;
    float support;
};
// imageops::blur: d_tmp and d_dst are h x w floats; d_dst may alias d_src's successor buffers but not d_src / d_tmp
cudaError_t blurb_gaussian(const float *d_src, int w, int h, const GaussTab &G, float *d_tmp, float *d_dst, cudaStream_t st);
// shi_tomasi_score: planes_a / planes_b are 3 x h x w floats of scratch, box_r the three box radii of fast_blur
cudaError_t scoreb_map(const float *d_img, int w, int h, const int box_r[3], float *d_planes_a, float *d_planes_b,
                       float *d_score, cudaStream_t st);
enum { DB_NCAND = 0, DB_NSURV, DB_MAXSCORE, DB_LASTROW, DB_N };
struct DetectB {
    const float *score;       // h x w
    int w, h;
    float threshold;
    int min_dist;
    float *cand;              // h x w scratch
    uint8_t *fmap;            // h x w scratch
    uint32_t *list, *surv;    // ceil(w/2) * ceil(h/2) entries each
    unsigned *cnt;            // DB_N counters, zero between frames
    unsigned long long *ids;  // feature table of the frame: [0, *n_feat) tracked features, new ones are appended
    float *xy;
    int *n_feat;              // device
    int n_tracked_host;       // upper bound of *n_feat known to the host (0: no tracked features)
    int cap;
    unsigned long long *last_id;  // device: next id
    int *n_new;               // device: number of appended features
    int *err;                 // device: set when the table overflowed
};
cudaError_t detectb_new_points(const DetectB &d, cudaStream_t st);
// tracked features of the new frame: status 0, old order, xy = T * xy_in
cudaError_t keepb_tracked(const unsigned long long *ids_in, const float *xy_in, int n, const float *T4, const int32_t *status,
                          unsigned long long *ids_out, float *xy_out, int *n_out, cudaStream_t st);

}  // namespace rsvio
#endif
