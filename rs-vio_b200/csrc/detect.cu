// detect.cu -- grid-based keypoint replenishment: FAST-9 per empty cell, weakest corner wins.
//
// Replaces detect_key_points (reference src/feature_tracker/image_utilities.rs:108-175) with
// num_points_in_cell == 1 (add_points, src/feature_tracker/feature_tracker.rs:222-251) and the
// imageproc 0.25 functions it calls: corners_fast9, is_corner_fast9, fast_corner_score
// (not vendored in the reference).
//
// What the reference does per empty cell: copy the g x g cell, then for thresholds 40,35,..,10
// run corners_fast9 (raster order), stable-sort ascending by score and keep the first corner
// that is at least 19 px from the image border.  The score is a binary search for the largest
// threshold at which the pixel is still a corner.  Restated in closed form:
//   score(p) = max over the 16 contiguous arcs of 9 circle pixels of
//              max(min_arc(I - c), min_arc(c - I)) - 1            (strict inequalities)
//   p is reported at threshold t iff score(p) >= t.
// With class(p) = min(40, 5 * floor(score / 5)) the pick is the candidate of the highest class
// with the smallest (score, raster index): one pass, one warp per cell, integer-only.
// Candidates are the cell pixels [3, g-3)^2 (the FAST circle must fit inside the cell copy).
#include "common.cuh"

namespace rsvio {

namespace {

constexpr int kEdge = 19;  // EDGE_THRESHOLD, image_utilities.rs:114

__device__ __forceinline__ int min9_max16(const int (&b)[16]) {
    // max over 16 cyclic windows of 9 of the window minimum, by doubling: 2, 4, 8, then +1
    int m2[16], m4[16], m8[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) m2[k] = min(b[k], b[(k + 1) & 15]);
#pragma unroll
    for (int k = 0; k < 16; ++k) m4[k] = min(m2[k], m2[(k + 2) & 15]);
#pragma unroll
    for (int k = 0; k < 16; ++k) m8[k] = min(m4[k], m4[(k + 4) & 15]);
    int best = -1000;
#pragma unroll
    for (int k = 0; k < 16; ++k) best = max(best, min(m8[k], b[(k + 8) & 15]));
    return best;
}

// quick reject at the lowest threshold used (10): a 9-arc always holds two adjacent compass points, so a
// corner needs two adjacent compass points both brighter than c + 10 or both darker than c - 10.
// p points at the pixel, s is the row stride.
__device__ __forceinline__ bool fast9_compass(const uint8_t *__restrict__ p, int s) {
    const int c = __ldg(p);
    const int p0 = __ldg(p - 3 * s), p4 = __ldg(p + 3), p8 = __ldg(p + 3 * s), p12 = __ldg(p - 3);
    const int hi = c + 10, lo = c - 10;
    const bool above = (p0 > hi && p4 > hi) || (p4 > hi && p8 > hi) || (p8 > hi && p12 > hi) || (p12 > hi && p0 > hi);
    const bool below = (p0 < lo && p4 < lo) || (p4 < lo && p8 < lo) || (p8 < lo && p12 < lo) || (p12 < lo && p0 < lo);
    return above || below;
}

// largest threshold at which (x, y) is a FAST-9 corner, < 10 if it is not one at threshold 10
__device__ __forceinline__ int fast9_score_full(const uint8_t *__restrict__ p, int s) {
    const int c = __ldg(p);
    int b[16];
    b[0] = __ldg(p - 3 * s);
    b[1] = __ldg(p - 3 * s + 1);
    b[2] = __ldg(p - 2 * s + 2);
    b[3] = __ldg(p - s + 3);
    b[4] = __ldg(p + 3);
    b[5] = __ldg(p + s + 3);
    b[6] = __ldg(p + 2 * s + 2);
    b[7] = __ldg(p + 3 * s + 1);
    b[8] = __ldg(p + 3 * s);
    b[9] = __ldg(p + 3 * s - 1);
    b[10] = __ldg(p + 2 * s - 2);
    b[11] = __ldg(p + s - 3);
    b[12] = __ldg(p - 3);
    b[13] = __ldg(p - s - 3);
    b[14] = __ldg(p - 2 * s - 2);
    b[15] = __ldg(p - 3 * s - 1);
    int d[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) {
        d[k] = c - b[k];
        b[k] = b[k] - c;
    }
    return max(min9_max16(b), min9_max16(d)) - 1;
}

// Occupied cells get "no point"; empty cells are appended (any order) to a global work list.
__global__ void __launch_bounds__(256)
list_empty_cells_kernel(GridDesc gd, int streams, const uint8_t *__restrict__ occ, uint32_t *__restrict__ cell_pt,
                        float *__restrict__ cell_score, uint32_t *__restrict__ list, int *__restrict__ list_count,
                        int *__restrict__ next_count, int *__restrict__ n_empty) {
    if (blockIdx.x == 0 && threadIdx.x == 0) *next_count = 0;  // the other counter, for the next frame
    const int cells = gd.ncx * gd.ncy;
    const long long total = (long long)cells * streams;
    const long long item = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    bool empty = false;
    if (item < total) {
        const int s = (int)(item / cells), cell = (int)(item - (long long)s * cells);
        const int cx = cell / gd.ncy, cy = cell - cx * gd.ncy;  // visiting order: x outer, y inner
        empty = __ldg(occ + (size_t)s * gd.occ_rows * gd.occ_cols + (size_t)cy * gd.occ_cols + cx) == 0;
        cell_pt[item] = 0xFFFFFFFFu;
        cell_score[item] = 0.0f;
    }
    const unsigned m = __ballot_sync(0xffffffffu, empty);
    const int lane = threadIdx.x & 31;
    int base = 0;
    if (lane == 0 && m) base = atomicAdd(list_count, __popc(m));
    base = __shfl_sync(0xffffffffu, base, 0);
    if (empty) list[base + __popc(m & ((1u << lane) - 1u))] = (uint32_t)item;
    if (empty) atomicAdd(n_empty + (int)(item / cells), 1);  // stateless path only (one image): not on the frame path
}

// CTA sizes of the detection kernel: 256 threads by default; 128 (64 registers x 128 threads = one CTA slot of a
// temporal LK launch with 7 resident CTAs per SM) when the launch has to run beside the next frame's LK.
constexpr int kDetChunk = 2048;  // candidates per compaction pass (a 40-px cell has 1156)

// ordered compaction of one stream's cells (visiting order: x outer, y inner) into identity transforms, by one CTA.
// The cell results were written by other CTAs of this launch (fenced before their `done` increment): read past L1.
template <int NT>
__device__ void collect_stream(const GridDesc &gd, int s, const uint32_t *cell_pt, const float *cell_score,
                               float *__restrict__ new_T6, float *__restrict__ new_score, int *__restrict__ new_count,
                               int new_cap, unsigned long long *counters, uint32_t *warp_tot /* >= 8 shared words */) {
    __shared__ int base;
    const int cells = gd.ncx * gd.ncy;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (threadIdx.x == 0) base = 0;
    __syncthreads();
    for (int c0 = 0; c0 < cells; c0 += NT) {
        const int c = c0 + threadIdx.x;
        const uint32_t pt = c < cells ? __ldcg(cell_pt + (size_t)s * cells + c) : 0xFFFFFFFFu;
        const bool has = pt != 0xFFFFFFFFu;
        const unsigned m = __ballot_sync(0xffffffffu, has);
        if (lane == 0) warp_tot[wid] = (uint32_t)__popc(m);
        __syncthreads();
        int off = base;
        for (int k = 0; k < wid; ++k) off += (int)warp_tot[k];
        const int rank = off + __popc(m & ((1u << lane) - 1u));
        if (has && rank < new_cap) {
            float *T = new_T6 + ((size_t)s * new_cap + rank) * 6;
            T[0] = 1.0f; T[1] = 0.0f; T[2] = (float)(pt & 0xFFFFu);
            T[3] = 0.0f; T[4] = 1.0f; T[5] = (float)(pt >> 16);
            if (new_score) new_score[(size_t)s * new_cap + rank] = __ldcg(cell_score + (size_t)s * cells + c);
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            int t = 0;
            for (int k = 0; k < NT / 32; ++k) t += (int)warp_tot[k];
            base += t;
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        const int n = base < new_cap ? base : new_cap;
        new_count[s] = n;
        if (counters) atomicAdd(counters + CNT_DETECTED, (unsigned long long)n);
    }
}

// One CTA per empty cell: the (g-6)^2 candidates are spread over 256 threads, block-wide packed-key max.
// (128 threads: 4 % slower on the bench step; one WARP per cell, no block barriers: 40 % slower -- a lane then
// tests 36 candidates in sequence.)
template <int NT>
__global__ void __launch_bounds__(NT)
detect_list_kernel(GridDesc gd, const uint8_t *__restrict__ img0, size_t stride, size_t pitch, CellWork cw,
                   float *__restrict__ new_T6, float *__restrict__ new_score, int new_cap, unsigned long long *counters) {
    const uint32_t *__restrict__ list = cw.list;
    uint32_t *__restrict__ cell_pt = cw.cell_pt;
    float *__restrict__ cell_score = cw.cell_score;
    __shared__ uint32_t wbest[NT / 32];
    __shared__ uint32_t surv[kDetChunk];
    __shared__ int n_surv;
    __shared__ int s_last;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int cells = gd.ncx * gd.ncy;
    const int g = (int)gd.g;
    const int side = g - 6;  // candidates per axis
    const int n = *cw.list_count;
    for (int i = blockIdx.x; i < n; i += gridDim.x) {
        const uint32_t item = __ldg(list + i);
        const int s = (int)(item / (uint32_t)cells), cell = (int)(item - (uint32_t)s * (uint32_t)cells);
        const int cx = cell / gd.ncy, cy = cell - cx * gd.ncy;
        const int x0 = (int)gd.x_start + cx * g, y0 = (int)gd.y_start + cy * g;
        const uint8_t *__restrict__ base = img0 + (size_t)s * pitch;
        uint32_t best = 0;
        // two passes per chunk of candidates: the cheap compass test compacts the survivors into a list, then the
        // 16-pixel score runs on full warps of survivors (a third of the pixels pass; testing and scoring in one
        // loop made every warp pay the full score in every round)
        for (int c0 = 0; c0 < side * side; c0 += kDetChunk) {
            if (threadIdx.x == 0) n_surv = 0;
            __syncthreads();
            const int c1 = min(c0 + kDetChunk, side * side);
            // tests first (independent loads of several candidates in flight), compaction afterwards
            uint32_t passbits = 0;
            static_assert(kDetChunk / NT <= 32, "one pass bit per candidate of a thread");
#pragma unroll 4
            for (int j = 0; j < kDetChunk / NT; ++j) {
                const int idx = c0 + j * NT + (int)threadIdx.x;
                if (idx < c1) {
                    const int ly = idx / side + 3, lx = idx - (idx / side) * side + 3;
                    const int gx = x0 + lx, gy = y0 + ly;
                    // point_in_bound(.., 19): inclusive upper bound (image_utilities.rs:68-73)
                    if (!(gx < kEdge || gy < kEdge || gx > gd.w - kEdge || gy > gd.h - kEdge) &&
                        fast9_compass(base + (size_t)gy * stride + gx, (int)stride))
                        passbits |= 1u << j;
                }
            }
            for (int j = 0; c0 + j * NT < c1; ++j) {
                const bool pass = (passbits >> j) & 1u;
                const unsigned m = __ballot_sync(0xffffffffu, pass);
                int at = 0;
                if (lane == 0 && m) at = atomicAdd(&n_surv, __popc(m));
                at = __shfl_sync(0xffffffffu, at, 0);
                if (pass) surv[at + __popc(m & ((1u << lane) - 1u))] = (uint32_t)(c0 + j * NT + (int)threadIdx.x);
            }
            __syncthreads();
            for (int k = threadIdx.x; k < n_surv; k += NT) {
                const int idx = (int)surv[k];
                const int ly = idx / side + 3, lx = idx - (idx / side) * side + 3;
                const int sc = fast9_score_full(base + (size_t)(y0 + ly) * stride + (x0 + lx), (int)stride);
                if (sc < 10) continue;
                const uint32_t cls = (uint32_t)(min(sc, 40) / 5);        // 2..8
                const uint32_t raster = (uint32_t)(ly * g + lx);        // order inside the cell copy
                best = max(best, (cls << 28) | ((uint32_t)(255 - sc) << 20) | (0xFFFFFu - raster));
            }
            __syncthreads();
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) best = max(best, __shfl_xor_sync(0xffffffffu, best, o));
        if (lane == 0) wbest[wid] = best;
        __syncthreads();
        if (threadIdx.x == 0) {
            for (int k = 1; k < NT / 32; ++k) best = max(best, wbest[k]);
            if (best != 0) {
                const int sc = 255 - (int)((best >> 20) & 0xFFu);
                const uint32_t raster = 0xFFFFFu - (best & 0xFFFFFu);
                const uint32_t ly = raster / (uint32_t)g, lx = raster - ly * (uint32_t)g;
                const uint32_t gx = gd.x_start + (uint32_t)cx * gd.g + lx, gy = gd.y_start + (uint32_t)cy * gd.g + ly;
                cell_pt[item] = (gy << 16) | gx;
                cell_score[item] = (float)sc;
            }
            // the CTA that finishes the last listed cell of a stream compacts that stream's new points (below)
            __threadfence();
            s_last = atomicAdd(cw.done + s, 1) + 1 == __ldg(cw.n_empty + s) ? 1 : 0;
        }
        __syncthreads();
        if (s_last) {
            __threadfence();
            collect_stream<NT>(gd, s, cell_pt, cell_score, new_T6, new_score, cw.new_count, new_cap, counters, wbest);
            if (threadIdx.x == 0) cw.done[s] = 0;  // ready for the next frame
            __syncthreads();
        }
    }
}

__global__ void mark_occupancy_xy_kernel(GridDesc gd, const uint32_t *__restrict__ xy, int n, uint8_t *occ) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t x = xy[2 * i], y = xy[2 * i + 1];
    const uint32_t x_end = gd.x_start + gd.g * (uint32_t)gd.ncx + 1, y_end = gd.y_start + gd.g * (uint32_t)gd.ncy + 1;
    if (x >= gd.x_start && y >= gd.y_start && x < x_end && y < y_end)
        occ[(size_t)((y - gd.y_start) / gd.g) * gd.occ_cols + (x - gd.x_start) / gd.g] = 1;
}

}  // namespace

cudaError_t detect_cells(const GridDesc &gd, const uint8_t *img0, size_t stride, size_t pitch, int streams,
                         const uint8_t *occ, const CellWork &cw, float *new_T6, float *new_score, int new_cap,
                         unsigned long long *counters, bool list_first, bool narrow, cudaStream_t st) {
    const long long total = (long long)gd.ncx * gd.ncy * streams;
    if (total <= 0) return cudaSuccess;
    if (list_first)  // stateless path: the frame path gets its list from the compaction kernel
        list_empty_cells_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(gd, streams, occ, cw.cell_pt, cw.cell_score,
                                                                                 cw.list, cw.list_count, cw.next_count,
                                                                                 cw.n_empty);
    // one resident wave of CTAs looping over the list (measured: 23 us; one CTA per candidate cell, most of
    // them exiting at once, 31 us)
    const long long wave = (long long)sm_count() * (narrow ? 16 : 8);  // the same number of threads either way
    long long blocks = total < wave ? total : wave;
    if (narrow)
        detect_list_kernel<128><<<(unsigned)blocks, 128, 0, st>>>(gd, img0, stride, pitch, cw, new_T6, new_score, new_cap, counters);
    else
        detect_list_kernel<256><<<(unsigned)blocks, 256, 0, st>>>(gd, img0, stride, pitch, cw, new_T6, new_score, new_cap, counters);
    return cudaGetLastError();
}

cudaError_t mark_occupancy_xy(const GridDesc &gd, const uint32_t *d_xy, int n, uint8_t *occ, cudaStream_t st) {
    if (n <= 0) return cudaSuccess;
    mark_occupancy_xy_kernel<<<(n + 255) / 256, 256, 0, st>>>(gd, d_xy, n, occ);
    return cudaGetLastError();
}

}  // namespace rsvio
