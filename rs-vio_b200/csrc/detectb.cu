// detectb.cu -- row f-2: the image side of the reference's second tracker (the `feature_tracker` crate, "crate B"):
// preprocessing blur, Shi-Tomasi score map, block non-maximum suppression, min-distance suppression against the
// tracked features, ordered list of new corners.
//
// Replaces (paths relative to the reference repository):
//   build_image_pyramid's `imageops::blur`                 feature_tracker/src/image_operations.rs:58-62
//   shi_tomasi_score                                        feature_tracker/src/feature_tracker/feature_detection.rs:83-165
//   suppress_non_maximum (radius 1 + threshold)             feature_detection.rs:172-254
//   add_points (tracked features injected with max + 1,     feature_detection.rs:48-80
//     `suppress::local_maxima(min_dist)`, border filter)
// Arithmetic that lives in crates the reference does not vendor is restated as oracle/crate_b.py documents it
// (image 0.25.9 `blur` / `fast_blur`, imageproc 0.26 `horizontal_filter` / `vertical_filter` / `local_maxima`):
// PARITY UNPINNED against the crates themselves, bit-identical to the oracle by construction -- every f32 operation
// is done in the oracle's order (the library is compiled with -fmad=false), including the sequential running sums
// of the box blur.
#include "common.cuh"

namespace rsvio {

namespace {

// ---- imageops::blur: vertical_sample then horizontal_sample with Filter{gaussian, support 2 sigma}, same size ----
// For a same-size pass ratio == sratio == 1 and the kernel argument (i - (o + 0.5 - 0.5)) is the integer i - o, so the
// un-normalised weights are one table g[d + K], d = i - o (made on the host with libm's expf, as Rust's f32::exp is).
__device__ __forceinline__ void gauss_window(int o, int n, float support, int &left, int &right) {
    const float c = ((float)o + 0.5f) * 1.0f;
    long long l = (long long)floorf(c - support), r = (long long)ceilf(c + support);
    l = l < 0 ? 0 : (l > (long long)n - 1 ? (long long)n - 1 : l);
    r = r < l + 1 ? l + 1 : (r > (long long)n ? (long long)n : r);
    left = (int)l;
    right = (int)r;
}

__global__ void gauss_vertical_kernel(const float *__restrict__ src, int w, int h, const __grid_constant__ GaussTab G,
                                      float *__restrict__ dst) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, o = blockIdx.y;
    if (x >= w) return;
    int left, right;
    gauss_window(o, h, G.support, left, right);
    float s = 0.0f;
    for (int i = left; i < right; ++i) s = s + G.g[i - o + G.K];
    float t = 0.0f;
    for (int i = left; i < right; ++i) t = t + src[(size_t)i * w + x] * (G.g[i - o + G.K] / s);
    dst[(size_t)o * w + x] = t;  // the intermediate is not clamped
}

__global__ void gauss_horizontal_kernel(const float *__restrict__ src, int w, int h, const __grid_constant__ GaussTab G,
                                        float *__restrict__ dst) {
    const int o = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (o >= w) return;
    int left, right;
    gauss_window(o, w, G.support, left, right);
    float s = 0.0f;
    for (int i = left; i < right; ++i) s = s + G.g[i - o + G.K];
    float t = 0.0f;
    for (int i = left; i < right; ++i) t = t + src[(size_t)y * w + i] * (G.g[i - o + G.K] / s);
    dst[(size_t)y * w + o] = fminf(fmaxf(t, 0.0f), 1.0f);
}

// ---- structure tensor: [-1, 0, 1] correlation (edges replicated), products --------------------------------------
__global__ void structure_kernel(const float *__restrict__ img, int w, int h, float *__restrict__ dxx,
                                 float *__restrict__ dyy, float *__restrict__ dxy) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= w) return;
    const float *__restrict__ r = img + (size_t)y * w;
    const float dx = r[min(x + 1, w - 1)] - r[max(x - 1, 0)];
    const float dy = img[(size_t)min(y + 1, h - 1) * w + x] - img[(size_t)max(y - 1, 0) * w + x];
    const size_t i = (size_t)y * w + x;
    dxx[i] = dx * dx;
    dyy[i] = dy * dy;
    dxy[i] = dx * dy;
}

// ---- fast_blur: one box pass = running sums along the rows, result / (2r + 1), clamp [0, 1], written TRANSPOSED ----
// The running sum `v = v - p[c - r] + p[c + r + 1]` is sequential by definition (its rounding errors travel along the
// row), so a row is one lane's serial chain.  A CTA owns 32 rows: warp 0 walks them, one row per lane, through a
// shared-memory tile of 64 columns (+ r columns of halo on each side); warps 1..3 stage the next tile meanwhile
// (row segments: coalesced reads, edges replicated), so the chain never waits for global memory, and the transposed
// writes are 128-byte lines.  (The first version had the one warp stage its own tiles: 0.4-0.6 ms per pass, 88 % of a
// crate B frame.)
constexpr int kBoxRMax = 40;  // sigma <= ~23
constexpr int kBoxChunk = 64;
constexpr int kBoxCols = kBoxChunk + 2 * kBoxRMax + 1;  // odd: lane * kBoxCols hits 32 different banks

__device__ __noinline__ float ieee_div(float a, float b) { return a / b; }

__global__ void __launch_bounds__(128)
box_pass_kernel(const float *__restrict__ src, int w, int h, int r, size_t plane, float *__restrict__ dst) {
    __shared__ float tile[2][32][kBoxCols];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, row0 = blockIdx.x * 32;
    src += (size_t)blockIdx.y * plane;
    dst += (size_t)blockIdx.y * plane;
    const int nrows = min(32, h - row0);
    const int span = kBoxChunk + 2 * r + 1;  // columns [c0 - r, c0 + 64 + r] of a chunk starting at c0
    const int nchunks = (w + kBoxChunk - 1) / kBoxChunk;
    // warps 1..3: tile[chunk & 1][rr][k] = src[row0 + rr][clamp(c0 - r + k)], as 4-byte cp.async copies -- all of a
    // thread's loads are in flight at once (with plain loads a warp fetched its 11 rows one after the other: a DRAM
    // round trip each, 13 us per chunk under ncu's cold caches)
    auto stage = [&](int chunk) {
        const int c0 = chunk * kBoxChunk;
        float(*t)[kBoxCols] = tile[chunk & 1];
        for (int rr = warp - 1; rr < nrows; rr += 3) {
            const float *__restrict__ p = src + (size_t)(row0 + rr) * w;
            for (int k = lane; k < span; k += 32) {
                const uint32_t sa = (uint32_t)__cvta_generic_to_shared(&t[rr][k]);
                asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(sa), "l"(p + min(max(c0 - r + k, 0), w - 1)) : "memory");
            }
        }
        asm volatile("cp.async.wait_all;" ::: "memory");
    };
    const int row = row0 + lane;
    const bool live = row < h;
    const float den = 2.0f * (float)r + 1.0f, rden = 1.0f / den;
    float v = 0.0f;
    if (warp == 0) {  // first window: columns -r .. r (clamped), summed left to right
        const float *__restrict__ p = src + (size_t)min(row, h - 1) * w;
        for (int x = -r; x <= r; ++x) v = v + p[min(max(x, 0), w - 1)];
    } else {
        stage(0);
    }
    __syncthreads();
    for (int chunk = 0; chunk < nchunks; ++chunk) {
        if (warp != 0) {
            if (chunk + 1 < nchunks) stage(chunk + 1);
        } else if (live) {
            const int c0 = chunk * kBoxChunk, nc = min(kBoxChunk, w - c0);
            const float *__restrict__ t = tile[chunk & 1][lane];
            // v / den, correctly rounded, without the IEEE division sequence (15 instructions with its range checks and
            // a slow-path call -- control flow that would keep the next step's shared-memory loads behind it):
            // q = RN(v y), r = v - q den (exact in the FMA), q' = RN(q + r y) equals RN(v / den) when y = RN(1 / den)
            // (Markstein) and nothing underflows.  Tiny non-zero running sums (impossible for f32 images in [0, 1],
            // whose squared differences are 0 or > 1e-15) are only flagged in the loop; a flagged chunk is redone with
            // the real division.
            const float v_start = v;
            bool tiny = false;
#pragma unroll 8  // the shared-memory loads run ahead of the two dependent adds per step
            for (int k = 0; k < nc; ++k) {
                float q = v * rden;
                q = fmaf(fmaf(-q, den, v), rden, q);
                tiny = tiny || (!(fabsf(v) > 1e-30f) && v != 0.0f);
                dst[(size_t)(c0 + k) * h + row] = fminf(fmaxf(q, 0.0f), 1.0f);
                v = v - t[k] + t[k + 2 * r + 1];  // p[c - r], p[c + r + 1]
            }
            if (tiny) {
                v = v_start;
                for (int k = 0; k < nc; ++k) {
                    dst[(size_t)(c0 + k) * h + row] = fminf(fmaxf(ieee_div(v, den), 0.0f), 1.0f);
                    v = v - t[k] + t[k + 2 * r + 1];
                }
            }
        }
        __syncthreads();
    }
}

__global__ void score_kernel(const float *__restrict__ dxx, const float *__restrict__ dyy, const float *__restrict__ dxy,
                             int n, float *__restrict__ score) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float a = dxx[i], b = dyy[i], c = dxy[i];
    const float trace = a + b;
    const float det = a * b - c * c;
    const float delta = fmaxf(trace * trace - 4.0f * det, 0.0f);
    score[i] = 500.0f * fabsf(trace - sqrtf(delta));
}

// ---- suppress_non_maximum, radius 1: one thread per 2x2 block ------------------------------------------------------
// cand[y][x] = score of a surviving block maximum, else -1 (scores are >= 0).  cnt[DB_MAXSCORE] = max score (as
// bits: non-negative floats order like unsigned), cnt[DB_LASTROW] = last row holding a candidate, list = their pixels.
__global__ void nms_kernel(const float *__restrict__ score, int w, int h, float threshold, float *__restrict__ cand,
                           uint32_t *__restrict__ list, unsigned *__restrict__ cnt) {
    const int bx0 = 2 * (blockIdx.x * blockDim.x + threadIdx.x), by0 = 2 * blockIdx.y;
    if (bx0 >= w || by0 >= h) return;
    int bx = bx0, by = by0;
    float bs = score[(size_t)by0 * w + bx0];
    const int y2 = min(h, by0 + 2), x2 = min(w, bx0 + 2);
    for (int cy = by0; cy < y2; ++cy)
        for (int cx = bx0; cx < x2; ++cx) {
            const float ci = score[(size_t)cy * w + cx];
            if (ci < bs) continue;
            if (ci > bs || (cx < bx || (cx == bx && cy < by))) {  // (cx, cy) < (best_x, best_y): x first
                bx = cx;
                by = cy;
                bs = ci;
            }
        }
    bool keep = bs >= threshold;
    if (keep) {
        const int x0 = max(bx - 1, 0), x3 = min(w, bx + 2), y0 = max(by - 1, 0), y3 = min(h, by + 2);
        for (int cy = y0; cy < y3 && keep; ++cy)
            for (int cx = x0; cx < x3; ++cx) {
                if (cy >= by0 && cy < y2 && cx >= bx0 && cx < x2) continue;  // the block itself
                const float ci = score[(size_t)cy * w + cx];
                if (ci < bs) continue;
                if (ci > bs || (cx < bx || (cx == bx && cy < by))) {
                    keep = false;
                    break;
                }
            }
    }
    for (int cy = by0; cy < y2; ++cy)
        for (int cx = bx0; cx < x2; ++cx) cand[(size_t)cy * w + cx] = (keep && cx == bx && cy == by) ? bs : -1.0f;
    if (keep) {
        list[atomicAdd(&cnt[DB_NCAND], 1u)] = ((uint32_t)by << 16) | (uint32_t)bx;
        atomicMax(&cnt[DB_MAXSCORE], __float_as_uint(bs));
        atomicMax(&cnt[DB_LASTROW], (unsigned)by);
    }
}

// tracked features: ShiTomasiCorner::from_feature rounds half away and casts to u32 (saturating)
__global__ void mark_features_kernel(const float2 *__restrict__ xy, const int *__restrict__ n_ptr, int w, int h,
                                     uint8_t *__restrict__ fmap, unsigned *__restrict__ cnt) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= *n_ptr) return;
    const uint32_t x = sat_u32(roundf(xy[i].x)), y = sat_u32(roundf(xy[i].y));
    atomicMax(&cnt[DB_LASTROW], y);  // `height` of local_maxima is the last occupied row, wherever it is
    if (x < (uint32_t)w && y < (uint32_t)h) fmap[(size_t)y * w + x] = 1;
}

// suppress::local_maxima over {candidates} + {tracked features, score max + 1}, then add_points' filters: keep the
// candidates that are no tracked feature's neighbour and have no better candidate in the (2r + 1)^2 block -- rows
// [cy - r, min(cy + r + 1, last_row)) only, as imageproc does.  One warp per candidate.
__global__ void local_maxima_kernel(const float *__restrict__ cand, const uint8_t *__restrict__ fmap,
                                    const uint32_t *__restrict__ list, unsigned *__restrict__ cnt, int w, int h, int r,
                                    int have_features, uint32_t *__restrict__ surv) {
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
    const int n = (int)cnt[DB_NCAND];
    const int last_row = (int)cnt[DB_LASTROW];
    for (int i = warp; i < n; i += nwarps) {
        const uint32_t key = list[i];
        const int cx = (int)(key & 0xffffu), cy = (int)(key >> 16);
        const float cs = cand[(size_t)cy * w + cx];
        const int row_lo = max(cy - r, 0), row_hi = min((cy + r + 1 > last_row) ? last_row : cy + r + 1, h);
        const int x_lo = max(cx - r, 0), x_hi = min(cx + r, w - 1);
        bool beaten = false;
        for (int y = row_lo; y < row_hi && !beaten; ++y) {
            bool b = false;
            for (int x = x_lo + lane; x <= x_hi; x += 32) {
                const float s = cand[(size_t)y * w + x];
                if (s > cs || (s == cs && (y < cy || (y == cy && x < cx)))) b = true;
                if (have_features && fmap[(size_t)y * w + x]) b = true;
            }
            beaten = __any_sync(0xffffffffu, b);
        }
        if (!beaten && lane == 0 && cx >= r && cx < w - r && cy >= r && cy < h - r)
            surv[atomicAdd(&cnt[DB_NSURV], 1u)] = key;
    }
}

// new features in (y, x) order behind the tracked ones: rank = number of survivors with a smaller key
__global__ void append_new_kernel(const uint32_t *__restrict__ surv, unsigned *__restrict__ cnt, int *__restrict__ n_feat,
                                  unsigned long long *__restrict__ ids, float2 *__restrict__ xy, int cap,
                                  unsigned long long *__restrict__ last_id, int *__restrict__ err) {
    const int n = (int)cnt[DB_NSURV], base = *n_feat;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const uint32_t key = surv[i];
        int rank = 0;
        for (int j = 0; j < n; ++j) rank += surv[j] < key ? 1 : 0;
        if (base + rank < cap) {
            ids[base + rank] = *last_id + (unsigned long long)rank;
            xy[base + rank] = make_float2((float)(key & 0xffffu), (float)(key >> 16));
        } else {
            *err = 1;
        }
    }
}
// after append_new_kernel (separate launch: every thread above reads the old values)
__global__ void finish_frame_kernel(unsigned *__restrict__ cnt, int *__restrict__ n_feat, int cap,
                                    unsigned long long *__restrict__ last_id, int *__restrict__ out_n_new) {
    const int n = (int)cnt[DB_NSURV];
    *last_id += (unsigned long long)n;  // next_id() ran once per new corner
    *out_n_new = n;
    *n_feat = min(*n_feat + n, cap);
    cnt[DB_NCAND] = cnt[DB_NSURV] = cnt[DB_MAXSCORE] = cnt[DB_LASTROW] = 0u;
}

// tracked features of the new frame: previous features whose forward-backward track succeeded, in their old order,
// moved by their transform (feature_tracker.rs:130-142).  One block; n is a few thousand at most.
__global__ void __launch_bounds__(1024)
keep_tracked_kernel(const unsigned long long *__restrict__ ids_in, const float2 *__restrict__ xy_in, int n,
                    const float4 *__restrict__ T, const int32_t *__restrict__ status,
                    unsigned long long *__restrict__ ids_out, float2 *__restrict__ xy_out, int *__restrict__ n_out) {
    __shared__ int warp_tot[32];
    __shared__ int base;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (threadIdx.x == 0) base = 0;
    __syncthreads();
    for (int i0 = 0; i0 < n; i0 += 1024) {
        const int i = i0 + threadIdx.x;
        const bool ok = i < n && status[i] == 0;
        const unsigned m = __ballot_sync(0xffffffffu, ok);
        if (lane == 0) warp_tot[wid] = __popc(m);
        __syncthreads();
        int off = base;
        for (int k = 0; k < wid; ++k) off += warp_tot[k];
        if (ok) {
            const int o = off + __popc(m & ((1u << lane) - 1u));
            const float4 t = T[i];
            const float2 p = xy_in[i];
            ids_out[o] = ids_in[i];
            xy_out[o] = make_float2(t.x * p.x - t.y * p.y + t.z, t.y * p.x + t.x * p.y + t.w);
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            int s = 0;
            for (int k = 0; k < 32; ++k) s += warp_tot[k];
            base += s;
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) *n_out = base;
}

}  // namespace

cudaError_t blurb_gaussian(const float *d_src, int w, int h, const GaussTab &G, float *d_tmp, float *d_dst, cudaStream_t st) {
    gauss_vertical_kernel<<<dim3((unsigned)((w + 127) / 128), (unsigned)h, 1), 128, 0, st>>>(d_src, w, h, G, d_tmp);
    gauss_horizontal_kernel<<<dim3((unsigned)((w + 127) / 128), (unsigned)h, 1), 128, 0, st>>>(d_tmp, w, h, G, d_dst);
    return cudaGetLastError();
}

cudaError_t scoreb_map(const float *d_img, int w, int h, const int box_r[3], float *d_planes_a, float *d_planes_b,
                       float *d_score, cudaStream_t st) {
    const size_t plane = (size_t)w * h;
    for (int k = 0; k < 3; ++k)
        if (box_r[k] < 0 || box_r[k] > kBoxRMax) return cudaErrorInvalidValue;
    structure_kernel<<<dim3((unsigned)((w + 127) / 128), (unsigned)h, 1), 128, 0, st>>>(d_img, w, h, d_planes_a,
                                                                                       d_planes_a + plane, d_planes_a + 2 * plane);
    for (int k = 0; k < 3; ++k) {  // a (h x w) -> b (w x h, transposed) -> a (h x w)
        box_pass_kernel<<<dim3((unsigned)((h + 31) / 32), 3, 1), 128, 0, st>>>(d_planes_a, w, h, box_r[k], plane, d_planes_b);
        box_pass_kernel<<<dim3((unsigned)((w + 31) / 32), 3, 1), 128, 0, st>>>(d_planes_b, h, w, box_r[k], plane, d_planes_a);
    }
    const int n = w * h;
    score_kernel<<<(n + 255) / 256, 256, 0, st>>>(d_planes_a, d_planes_a + plane, d_planes_a + 2 * plane, n, d_score);
    return cudaGetLastError();
}

cudaError_t detectb_new_points(const DetectB &d, cudaStream_t st) {
    cudaError_t e = cudaMemsetAsync(d.fmap, 0, (size_t)d.w * d.h, st);
    if (e != cudaSuccess) return e;
    nms_kernel<<<dim3((unsigned)(((d.w + 1) / 2 + 127) / 128), (unsigned)((d.h + 1) / 2), 1), 128, 0, st>>>(
        d.score, d.w, d.h, d.threshold, d.cand, d.list, d.cnt);
    if (d.n_tracked_host > 0)
        mark_features_kernel<<<(d.n_tracked_host + 127) / 128, 128, 0, st>>>(reinterpret_cast<const float2 *>(d.xy), d.n_feat, d.w, d.h, d.fmap, d.cnt);
    local_maxima_kernel<<<sm_count() * 4, 256, 0, st>>>(d.cand, d.fmap, d.list, d.cnt, d.w, d.h, d.min_dist,
                                                          d.n_tracked_host > 0 ? 1 : 0, d.surv);
    append_new_kernel<<<sm_count(), 256, 0, st>>>(d.surv, d.cnt, d.n_feat, d.ids, reinterpret_cast<float2 *>(d.xy), d.cap, d.last_id, d.err);
    finish_frame_kernel<<<1, 1, 0, st>>>(d.cnt, d.n_feat, d.cap, d.last_id, d.n_new);
    return cudaGetLastError();
}

cudaError_t keepb_tracked(const unsigned long long *ids_in, const float *xy_in, int n, const float *T4, const int32_t *status,
                          unsigned long long *ids_out, float *xy_out, int *n_out, cudaStream_t st) {
    keep_tracked_kernel<<<1, 1024, 0, st>>>(ids_in, reinterpret_cast<const float2 *>(xy_in), n,
                                            reinterpret_cast<const float4 *>(T4), status, ids_out,
                                            reinterpret_cast<float2 *>(xy_out), n_out);
    return cudaGetLastError();
}

}  // namespace rsvio
