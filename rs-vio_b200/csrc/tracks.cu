// tracks.cu -- device-resident track tables: survivor compaction, occupancy grid, id assignment,
// packed output.
//
// Replaces the HashMap bookkeeping of StereoPatchTracker::process_frame
// (reference src/feature_tracker/feature_tracker.rs:116-187):
//   :125-138  tracked_points_map_cam{0,1} = track_points(..)   -> compact_tracks_kernel
//   :228-239  add_points: rounded tracked positions -> occupancy (image_utilities.rs:128-139)
//   :143-170  tmp maps, stereo match, `last_keypoint_id += 1` -> merge_new_kernel
//   :177-186  frame.add_left_feature / add_right_feature      -> packed (id, x, y) records
//   :201-206  remove_id                                        -> remove_ids_kernel
// Tables stay sorted by id: survivors keep their order, new ids are larger than all old ones and
// are handed out in ascending detection order (the reference uses HashMap iteration order, which
// is randomly seeded, so ids are only defined up to relabelling there).
#include "common.cuh"

namespace rsvio {

namespace {

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;

// block-wide stable compaction helper: returns the output rank of this thread's element
// (valid only when keep), advances *base by the number kept in this chunk.
template <int NW = kWarps>
__device__ __forceinline__ int block_rank(bool keep, int *warp_tot, int *base) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const unsigned m = __ballot_sync(0xffffffffu, keep);
    if (lane == 0) warp_tot[wid] = __popc(m);
    __syncthreads();
    int off = *base;
    for (int k = 0; k < wid; ++k) off += warp_tot[k];
    const int rank = off + __popc(m & ((1u << lane) - 1u));
    __syncthreads();
    if (threadIdx.x == 0) {
        int t = 0;
        for (int k = 0; k < NW; ++k) t += warp_tot[k];
        *base += t;
    }
    __syncthreads();
    return rank;
}

// block-wide exclusive offset of a per-thread count (0 .. E); advances *base by the block total
template <int NW>
__device__ __forceinline__ int block_offset(int cnt, int *warp_tot, int *base) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    int incl = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) warp_tot[wid] = incl;
    __syncthreads();
    int off = *base + incl - cnt;
    for (int k = 0; k < wid; ++k) off += warp_tot[k];
    __syncthreads();
    if (threadIdx.x == 0) {
        int t = 0;
        for (int k = 0; k < NW; ++k) t += warp_tot[k];
        *base += t;
    }
    __syncthreads();
    return off;
}

// Tables of up to kThreads * 4 tracks per group (the usual few hundred tracks of a stream): one CTA per group walks the
// table, builds the stream's occupancy grid and lists its empty cells -- the fewest dependent memory round trips.
__global__ void __launch_bounds__(kThreads)
compact_tracks_small_kernel(TableView in, const float *__restrict__ lk_T6, const int32_t *__restrict__ lk_status,
                      TableView out, int cap, int streams, GridDesc gd, uint8_t *__restrict__ occ, int have_prev,
                      unsigned long long *counters, int *__restrict__ err_clear, CellWork cw, int *__restrict__ surv) {
    // the overflow flag the merge kernel of this frame may raise (its previous use, two frames ago, has been read)
    if (err_clear && blockIdx.x == 0 && threadIdx.x == 0) *err_clear = 0;
    __shared__ int warp_tot[kWarps];
    __shared__ int base;
    const int g = blockIdx.x;
    const int n = in.counts[g];
    const bool cam0 = g < streams;
    uint8_t *my_occ = occ + (size_t)g * gd.occ_rows * gd.occ_cols;  // only used when cam0
    if (threadIdx.x == 0) base = 0;
    if (cam0)
        for (int i = threadIdx.x; i < gd.occ_rows * gd.occ_cols; i += kThreads) my_occ[i] = 0;
    __syncthreads();
    const uint32_t x_end = gd.x_start + gd.g * (uint32_t)gd.ncx + 1, y_end = gd.y_start + gd.g * (uint32_t)gd.ncy + 1;
    for (int c0 = 0; c0 < n; c0 += kThreads) {
        const int k = c0 + threadIdx.x;
        const size_t src = (size_t)g * cap + k;
        bool keep = false;
        float T[6];
        if (k < n) {
            // first frame after create/reset: nothing was tracked, the table is carried over as is
            keep = have_prev ? (lk_status[src] == RSVIO_TRACK_OK) : true;
            const float *ps = have_prev ? lk_T6 + src * 6 : in.T6 + src * 6;
#pragma unroll
            for (int q = 0; q < 6; ++q) T[q] = ps[q];
        }
        const int rank = block_rank(keep, warp_tot, &base);
        if (keep) {
            const size_t dst = (size_t)g * cap + rank;
            out.ids[dst] = in.ids[src];
#pragma unroll
            for (int q = 0; q < 6; ++q) out.T6[dst * 6 + q] = T[q];
            if (cam0) {  // add_points: Corner::new(round(x) as u32, round(y) as u32)
                const uint32_t x = sat_u32(roundf(T[2])), y = sat_u32(roundf(T[5]));
                if (x >= gd.x_start && y >= gd.y_start && x < x_end && y < y_end)
                    my_occ[(size_t)((y - gd.y_start) / gd.g) * gd.occ_cols + (x - gd.x_start) / gd.g] = 1;
            }
        }
    }
    if (threadIdx.x == 0) {
        out.counts[g] = base;
        if (surv) surv[g] = base;
        if (counters && g == 0) atomicAdd(counters + CNT_FRAMES, (unsigned long long)streams);
    }
    // The work list of the detection kernel (detect_key_points visits the cells without a tracked point,
    // image_utilities.rs:141-149): the CTA that built a stream's occupancy grid appends that stream's empty cells
    // (any order: results are per cell) and initialises every cell to "no point".
    if (cw.list && cam0) {
        if (g == 0 && threadIdx.x == 0) *cw.next_count = 0;  // the other counter of the pair, for the next frame
        __syncthreads();  // occupancy complete (written by this CTA)
        const int cells = gd.ncx * gd.ncy;
        const int lane = threadIdx.x & 31;
        int n_empty = 0;
        for (int c0 = 0; c0 < cells; c0 += kThreads) {
            const int cell = c0 + threadIdx.x;
            bool empty = false;
            if (cell < cells) {
                const int cx = cell / gd.ncy, cy = cell - cx * gd.ncy;  // visiting order: x outer, y inner
                empty = my_occ[(size_t)cy * gd.occ_cols + cx] == 0;
                cw.cell_pt[(size_t)g * cells + cell] = 0xFFFFFFFFu;
            }
            const unsigned m = __ballot_sync(0xffffffffu, empty);
            int at = 0;
            if (lane == 0 && m) at = atomicAdd(cw.list_count, __popc(m));
            at = __shfl_sync(0xffffffffu, at, 0);
            if (empty) cw.list[at + __popc(m & ((1u << lane) - 1u))] = (uint32_t)((size_t)g * cells + cell);
            n_empty += __syncthreads_count(empty);
        }
        if (threadIdx.x == 0) {
            cw.n_empty[g] = n_empty;
            if (n_empty == 0) cw.new_count[g] = 0;  // nobody will collect this stream
        }
    }
}


// One CTA per chunk of kChunk consecutive tracks of a group (grid: chunks x groups).  The usual few hundred tracks of a
// stream are one chunk; a high-resolution stream (10 240 cells keep ~13 k tracks per camera) spreads over a dozen SMs
// instead of walking the table with one CTA (90 us of a 690 us frame, the SM's load/store queue the limiter:
// lg_throttle 36 warps per issue).  A CTA finds its output offset by counting the survivors of the chunks before it
// (coalesced int4 reads of the status array), ranks its own chunk, and then moves the rows as consecutive floats of
// the OUTPUT (fully coalesced stores, near-coalesced loads: survivors are mostly neighbours).  The CTA that arrives
// last at a camera-0 group (per-group counter, reset by that CTA) lists the stream's empty cells and clears the
// occupancy grid for the next frame.
constexpr int kCompactThreads = 256, kCompactE = 4, kChunk = kCompactThreads * kCompactE;
__global__ void __launch_bounds__(kCompactThreads)
compact_tracks_kernel(TableView in, const float *__restrict__ lk_T6, const int32_t *__restrict__ lk_status,
                      TableView out, int cap, int streams, GridDesc gd, uint8_t *__restrict__ occ, int have_prev,
                      unsigned long long *counters, int *__restrict__ err_clear, CellWork cw, int *__restrict__ surv) {
    constexpr int NT = kCompactThreads, E = kCompactE, NW = NT / 32;
    const int g = blockIdx.y, chunk = blockIdx.x, n_chunks = gridDim.x;
    const int tid = threadIdx.x;
    // the overflow flag the merge kernel of this frame may raise (its previous use, two frames ago, has been read)
    if (err_clear && g == 0 && chunk == 0 && tid == 0) *err_clear = 0;
    __shared__ int s_src[kChunk];  // rank within the chunk -> index within the chunk
    __shared__ int warp_tot[NW];
    __shared__ int s_base;
    __shared__ int s_last;
    const int n = in.counts[g];
    const int start = chunk * kChunk;
    const bool cam0 = g < streams;
    uint8_t *my_occ = occ + (size_t)g * gd.occ_rows * gd.occ_cols;  // only used when cam0
    const size_t row0 = (size_t)g * cap;
    if (tid == 0) s_base = 0;
    __syncthreads();
    if (start < n) {
        // survivors of the chunks before this one (first frame after create/reset: nothing was tracked, the table is
        // carried over as is)
        int before = start;
        if (have_prev && start > 0) {
            int c = 0;  // row0 and start are multiples of 4 (cap = 4 cells + 64): aligned int4 reads
            const int4 *__restrict__ sp = reinterpret_cast<const int4 *>(lk_status + row0);
#pragma unroll 4
            for (int i = tid; i < start / 4; i += NT) {
                const int4 q = __ldg(sp + i);
                c += (q.x == RSVIO_TRACK_OK) + (q.y == RSVIO_TRACK_OK) + (q.z == RSVIO_TRACK_OK) + (q.w == RSVIO_TRACK_OK);
            }
            block_offset<NW>(c, warp_tot, &s_base);
            before = s_base;
            __syncthreads();
            if (tid == 0) s_base = 0;
            __syncthreads();
        }
        const int m = min(kChunk, n - start);
        bool keep[E];
        int cnt = 0;
#pragma unroll
        for (int e = 0; e < E; ++e) {
            const int k = tid * E + e;
            keep[e] = k < m && (have_prev ? (lk_status[row0 + start + k] == RSVIO_TRACK_OK) : true);
            cnt += keep[e] ? 1 : 0;
        }
        int rank = block_offset<NW>(cnt, warp_tot, &s_base);
#pragma unroll
        for (int e = 0; e < E; ++e)
            if (keep[e]) s_src[rank++] = tid * E + e;
        __syncthreads();
        const int kept = s_base;
        const float *__restrict__ srcT = (have_prev ? lk_T6 : in.T6) + (row0 + start) * 6;
        float *__restrict__ dstT = out.T6 + (row0 + before) * 6;
        RSVIO_CHECK(kept >= 0 && kept <= m && before >= 0 && before + kept <= cap);
#pragma unroll 8  // independent gathers: 8 in flight per thread
        for (int f = tid; f < kept * 6; f += NT) {
            const int r = f / 6, q = f - r * 6;
            dstT[f] = srcT[s_src[r] * 6 + q];
        }
        const uint32_t x_end = gd.x_start + gd.g * (uint32_t)gd.ncx + 1, y_end = gd.y_start + gd.g * (uint32_t)gd.ncy + 1;
#pragma unroll 2
        for (int r = tid; r < kept; r += NT) {
            const int k = s_src[r];
            RSVIO_CHECK(k >= 0 && k < m);
            out.ids[row0 + before + r] = in.ids[row0 + start + k];
            if (cam0) {  // add_points: Corner::new(round(x) as u32, round(y) as u32)
                const uint32_t x = sat_u32(roundf(srcT[k * 6 + 2])), y = sat_u32(roundf(srcT[k * 6 + 5]));
                if (x >= gd.x_start && y >= gd.y_start && x < x_end && y < y_end)
                    my_occ[(size_t)((y - gd.y_start) / gd.g) * gd.occ_cols + (x - gd.x_start) / gd.g] = 1;
            }
        }
        if (tid == 0 && start + kChunk >= n) {  // the group's last chunk knows the total
            out.counts[g] = before + kept;
            if (surv) surv[g] = before + kept;
        }
    } else if (chunk == 0 && tid == 0) {
        out.counts[g] = 0;
        if (surv) surv[g] = 0;
    }
    if (tid == 0 && counters && g == 0 && chunk == 0) atomicAdd(counters + CNT_FRAMES, (unsigned long long)streams);
    // The work list of the detection kernel (detect_key_points visits the cells without a tracked point,
    // image_utilities.rs:141-149): the CTA that completes a stream's occupancy grid appends that stream's empty cells
    // (any order: results are per cell) and initialises every cell to "no point".
    if (cw.list && cam0) {
        if (g == 0 && chunk == 0 && tid == 0) *cw.next_count = 0;  // the other counter of the pair, for the next frame
        __threadfence();  // this CTA's occupancy marks are visible device-wide before it arrives
        __syncthreads();
        if (n_chunks > 1) {
            if (tid == 0) {
                s_last = atomicAdd(cw.arrive + g, 1) == n_chunks - 1;
                if (s_last) cw.arrive[g] = 0;  // every chunk has arrived: ready for the next frame
            }
            __syncthreads();
            if (!s_last) return;
            __threadfence();
        }
        // every thread takes a run of consecutive cells: count its empty ones, ONE reservation in the list for the whole
        // CTA (a reservation per warp and round was 40 dependent atomics for 10 240 cells), then append
        const int cells = gd.ncx * gd.ncy;
        const int per = (cells + NT - 1) / NT;
        const int c_lo = min(tid * per, cells), c_hi = min(c_lo + per, cells);
        int mine = 0;
#pragma unroll 8
        for (int cell = c_lo; cell < c_hi; ++cell) {
            const int cx = cell / gd.ncy, cy = cell - cx * gd.ncy;  // visiting order: x outer, y inner
            mine += __ldcg(my_occ + (size_t)cy * gd.occ_cols + cx) == 0;  // L2: other CTAs wrote it
            cw.cell_pt[(size_t)g * cells + cell] = 0xFFFFFFFFu;
        }
        if (tid == 0) s_base = 0;
        __syncthreads();
        int at = block_offset<NW>(mine, warp_tot, &s_base);
        const int n_empty = s_base;
        __syncthreads();
        if (tid == 0) {
            s_base = n_empty ? atomicAdd(cw.list_count, n_empty) : 0;
            cw.n_empty[g] = n_empty;
            if (n_empty == 0) cw.new_count[g] = 0;  // nobody will collect this stream
        }
        __syncthreads();
        at += s_base;
        for (int cell = c_lo; cell < c_hi; ++cell) {
            const int cx = cell / gd.ncy, cy = cell - cx * gd.ncy;
            if (__ldcg(my_occ + (size_t)cy * gd.occ_cols + cx) == 0) cw.list[at++] = (uint32_t)((size_t)g * cells + cell);
        }
        __syncthreads();
        // every read of the grid is behind the barrier: clear it for the next frame's marks
        for (int i = tid; i < gd.occ_rows * gd.occ_cols; i += NT) my_occ[i] = 0;
    }
}

__global__ void __launch_bounds__(kThreads)
merge_new_kernel(TableView tab, int streams, int cap, int stereo, const float *__restrict__ new_T6,
                 const float *__restrict__ st_T6, const int32_t *__restrict__ st_status,
                 const int *__restrict__ new_count, int new_cap, unsigned long long *__restrict__ last_id,
                 int *err_flag) {
    __shared__ int warp_tot[kWarps];
    __shared__ int base;
    const int s = blockIdx.x;
    const int n_new = new_count[s];
    const int g0 = s, g1 = streams + s;
    const int n0 = tab.counts[g0];
    const int n1 = stereo ? tab.counts[g1] : 0;
    const unsigned long long id0 = last_id[s];
    if (threadIdx.x == 0) base = 0;
    __syncthreads();
    for (int c0 = 0; c0 < n_new; c0 += kThreads) {
        const int k = c0 + threadIdx.x;
        const size_t src = (size_t)s * new_cap + k;
        const bool keep = k < n_new && (!stereo || st_status[src] == RSVIO_TRACK_OK);
        const int rank = block_rank(keep, warp_tot, &base);
        if (keep) {
            if (n0 + rank < cap && n1 + rank < cap) {
                const size_t d0 = (size_t)g0 * cap + n0 + rank;
                tab.ids[d0] = id0 + (unsigned long long)rank;
#pragma unroll
                for (int q = 0; q < 6; ++q) tab.T6[d0 * 6 + q] = new_T6[src * 6 + q];
                if (stereo) {
                    const size_t d1 = (size_t)g1 * cap + n1 + rank;
                    tab.ids[d1] = id0 + (unsigned long long)rank;
#pragma unroll
                    for (int q = 0; q < 6; ++q) tab.T6[d1 * 6 + q] = st_T6[src * 6 + q];
                }
            } else {
                atomicExch(err_flag, 1);
            }
        }
    }
    __syncthreads();
    int add = base;
    if (n0 + add > cap) add = cap - n0;
    if (stereo && n1 + add > cap) add = cap - n1 < add ? cap - n1 : add;
    if (add < 0) add = 0;
    const int f0 = n0 + add, f1 = stereo ? n1 + add : 0;
    __syncthreads();
    if (threadIdx.x == 0) {
        tab.counts[g0] = f0;
        if (stereo) tab.counts[g1] = f1;
        last_id[s] = id0 + (unsigned long long)add;
    }
}

// packed output (mapped host memory): get_track_points / frame population (feature_tracker.rs:177-200).
// Own kernel so that it can run beside the next frame: only live records cross PCIe.
// 128-thread CTAs (x 8 chunks per group): 56 registers x 128 threads fit the one CTA slot per SM that a temporal LK
// launch with 7 resident CTAs leaves free, so the export of frame t runs beside the LK launch of frame t + 1
constexpr int kExportChunks = 8, kExportThreads = 128;
__global__ void __launch_bounds__(kExportThreads)
export_records_kernel(TableView tab, int cap, int streams, OutRecord *__restrict__ out_rec, int *__restrict__ out_counts,
                      const __grid_constant__ CameraModelDev cam0, const __grid_constant__ CameraModelDev cam1,
                      float2 *__restrict__ out_uxy, const int *__restrict__ err_flag) {
    const int g = blockIdx.x;
    const int n = tab.counts[g];
    if (blockIdx.y == 0 && threadIdx.x == 0) out_counts[g] = n;
    // the frame's overflow flag rides along behind the counts: no memcpy on the per-frame path at all
    if (g == 0 && blockIdx.y == 0 && threadIdx.x == 0 && err_flag) out_counts[gridDim.x] = *err_flag;
    const CameraModelDev &cm = g < streams ? cam0 : cam1;
    const bool unproject = out_uxy != nullptr && cm.model != RSVIO_CAMERA_NONE;
    for (int k = blockIdx.y * kExportThreads + threadIdx.x; k < n; k += kExportChunks * kExportThreads) {
        const size_t i = (size_t)g * cap + k;
        OutRecord r;
        r.id = tab.ids[i];
        r.x = tab.T6[i * 6 + 2];
        r.y = tab.T6[i * 6 + 5];
        out_rec[i] = r;
        if (unproject) {  // Frame::add_left_feature / add_right_feature (estimator/frame.rs:118-119,131-132)
            double ux, uy;
            unproject_pixel(cm, (double)r.x, (double)r.y, ux, uy);
            out_uxy[i] = make_float2((float)ux, (float)uy);
        }
    }
}

__global__ void unproject_points_kernel(const __grid_constant__ CameraModelDev cm, const float2 *__restrict__ xy, int n,
                                        float2 *__restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double ux, uy;
    unproject_pixel(cm, (double)xy[i].x, (double)xy[i].y, ux, uy);
    out[i] = make_float2((float)ux, (float)uy);
}

__global__ void __launch_bounds__(kThreads)
remove_ids_kernel(TableView in, TableView out, int cap, int streams, int stream,
                  const unsigned long long *__restrict__ rm, int n_rm) {
    __shared__ int warp_tot[kWarps];
    __shared__ int base;
    const int g = blockIdx.x;
    const int n = in.counts[g];
    const bool affected = (g % streams) == stream;
    if (threadIdx.x == 0) base = 0;
    __syncthreads();
    for (int c0 = 0; c0 < n; c0 += kThreads) {
        const int k = c0 + threadIdx.x;
        const size_t src = (size_t)g * cap + k;
        bool keep = k < n;
        unsigned long long id = 0;
        if (keep) {
            id = in.ids[src];
            if (affected)
                for (int j = 0; j < n_rm; ++j) keep = keep && (rm[j] != id);
        }
        const int rank = block_rank(keep, warp_tot, &base);
        if (keep) {
            const size_t dst = (size_t)g * cap + rank;
            out.ids[dst] = id;
#pragma unroll
            for (int q = 0; q < 6; ++q) out.T6[dst * 6 + q] = in.T6[src * 6 + q];
        }
    }
    if (threadIdx.x == 0) out.counts[g] = base;
}

}  // namespace

cudaError_t compact_tracks(const TableView &in, const float *lk_T6, const int32_t *lk_status, const TableView &out,
                           int groups, int cap, int streams, const GridDesc &gd, uint8_t *occ, int have_prev,
                           unsigned long long *counters, int *err_clear, const CellWork &cw, int *surv, cudaStream_t st) {
    if (cap <= kChunk) {
        compact_tracks_small_kernel<<<groups, kThreads, 0, st>>>(in, lk_T6, lk_status, out, cap, streams, gd, occ, have_prev,
                                                               counters, err_clear, cw, surv);
    } else {
        const int chunks = (cap + kChunk - 1) / kChunk;
        compact_tracks_kernel<<<dim3((unsigned)chunks, (unsigned)groups, 1), kCompactThreads, 0, st>>>(
            in, lk_T6, lk_status, out, cap, streams, gd, occ, have_prev, counters, err_clear, cw, surv);
    }
    return cudaGetLastError();
}

cudaError_t merge_new_points(const TableView &tab, int streams, int cap, int stereo, const float *new_T6,
                             const float *st_T6, const int32_t *st_status, const int *new_count, int new_cap,
                             unsigned long long *last_id, int *err_flag, cudaStream_t st) {
    merge_new_kernel<<<streams, kThreads, 0, st>>>(tab, streams, cap, stereo, new_T6, st_T6, st_status, new_count,
                                                   new_cap, last_id, err_flag);
    return cudaGetLastError();
}

cudaError_t export_records(const TableView &tab, int groups, int cap, int streams, OutRecord *out_rec, int *out_counts,
                           const CameraModelDev &cam0, const CameraModelDev &cam1, float *out_uxy, const int *err_flag,
                           cudaStream_t st) {
    export_records_kernel<<<dim3((unsigned)groups, kExportChunks, 1), kExportThreads, 0, st>>>(
        tab, cap, streams, out_rec, out_counts, cam0, cam1, reinterpret_cast<float2 *>(out_uxy), err_flag);
    return cudaGetLastError();
}

cudaError_t unproject_points(const CameraModelDev &cm, const float *d_xy, int n, float *d_out, cudaStream_t st) {
    if (n <= 0) return cudaSuccess;
    unproject_points_kernel<<<(n + 255) / 256, 256, 0, st>>>(cm, reinterpret_cast<const float2 *>(d_xy), n,
                                                             reinterpret_cast<float2 *>(d_out));
    return cudaGetLastError();
}

cudaError_t remove_ids(const TableView &in, const TableView &out, int groups, int cap, int streams, int stream,
                       const unsigned long long *d_ids, int n, cudaStream_t st) {
    remove_ids_kernel<<<groups, kThreads, 0, st>>>(in, out, cap, streams, stream, d_ids, n);
    return cudaGetLastError();
}

}  // namespace rsvio
