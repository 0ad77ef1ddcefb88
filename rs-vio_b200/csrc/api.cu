// api.cu -- the extern "C" boundary declared in include/rsvio_b200.h.
//
// Mirrors the reference's StereoPatchTracker<N> / PatchTracker<N>
// (src/feature_tracker/feature_tracker.rs:36-207) as an opaque handle that owns all device
// state: two ping-pong sets of pyramids (previous / current), the per-camera track tables, the
// detection scratch and a host-mapped output ring.  One frame is one stream-ordered sequence
//   upload -> pyramid -> temporal LK (both cameras, all streams, one launch) -> compaction +
//   occupancy -> FAST detection -> collect -> stereo LK -> merge + packed output
// with no host round trip in between.  There is no CPU fallback anywhere in this file.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <new>
#include <string>
#include <vector>

#include "common.cuh"

using namespace rsvio;

namespace {

constexpr int kSlots = 2;

// Every entry point runs on the handle's (or the requested) device and leaves the caller's current device as
// it found it.
struct DeviceScope {
    int prev = -1, cur = -1;
    cudaError_t enter(int dev) {
        if (cudaGetDevice(&prev) != cudaSuccess) {
            cudaGetLastError();
            prev = -1;
        }
        cur = dev;
        return prev == dev ? cudaSuccess : cudaSetDevice(dev);
    }
    ~DeviceScope() {
        if (prev >= 0 && prev != cur) cudaSetDevice(prev);
    }
};

#define CU_TRY(t, expr)                                                                        \
    do {                                                                                       \
        cudaError_t e__ = (expr);                                                              \
        if (e__ != cudaSuccess) return fail_cuda((t), e__, #expr, __LINE__);                   \
    } while (0)

}  // namespace

struct rsvio_tracker {
    rsvio_config cfg{};
    int S = 1, groups = 2, cap = 0, new_cap = 0, stereo = 1;
    size_t img_bytes = 0;
    cudaStream_t st = nullptr, copy_st = nullptr;
    cudaStream_t exp_st = nullptr;  // record export + error flag: beside the next frame on big batches, else == st
    bool own_exp = false;
    cudaStream_t pyr_st = nullptr;  // pyramid of frame t+1 beside detection / stereo LK / merge of frame t (big batches), else == st
    bool own_pyr = false;
    cudaEvent_t ev_lk_done = nullptr, ev_pyr[kSlots]{};
    // Batched handles split the temporal LK launch of a frame: the survivors of the previous frame's compaction (the
    // bulk) are tracked on lk_st as soon as that compaction and this frame's pyramid are done -- beside the previous
    // frame's detection / stereo LK / merge -- and only the points that merge appended wait for it (on st).
    cudaStream_t lk_st = nullptr;
    cudaEvent_t ev_compact = nullptr, ev_lk1 = nullptr;
    int *d_surv = nullptr;     // [groups] survivors per group after the last compaction (rows before the merged new points)
    bool split_lk = false;     // the handle may split (batched, own streams)
    bool surv_valid = false;   // d_surv describes tab[tcur] (false after reset / remove_ids)
    bool last_split = false;   // the frame just enqueued used the split launch
    bool lk_helper = true;     // split launch: a third launch fills the free CTA slot once the previous frame's tail is done
    // Host-fed frames arrive at PCIe rate, about half the rate the kernels could take them at EuRoC size: the device idles
    // between frames, there is no tail to hide, and the split only adds latency (7 instead of 8 CTAs per SM at the start,
    // narrower tail kernels) -- which matters with two frames in flight and a host loop that has to enqueue the next
    // upload in time.  sync() looks whether the next frame's upload was still running when this frame finished; while
    // that is so, frames are enqueued unsplit.
    bool upload_bound = false;
    bool slot_host[kSlots] = {false, false};  // the frame in this slot came through upload()
    bool own_stream = false;
    PyrPlan *plan = nullptr;
    PyrDesc pd{};
    GridDesc gd{};
    uint8_t *d_in[kSlots] = {nullptr, nullptr};
    uint8_t *h_stage[kSlots] = {nullptr, nullptr};
    uint8_t *d_pyr[2] = {nullptr, nullptr};
    int pcur = 0;  // index of the set that will receive the next frame's pyramids
    bool have_prev = false;
    TableView tab[2]{};
    int tcur = 0;  // table holding the current tracks
    float *d_lk_T6 = nullptr;
    int32_t *d_lk_status = nullptr;
    uint8_t *d_occ = nullptr;
    uint32_t *d_cell_pt = nullptr;
    float *d_cell_score = nullptr;
    uint32_t *d_cell_list = nullptr;
    int *d_cell_list_n = nullptr;
    int *d_cell_work = nullptr;  // [0, S): listed empty cells per stream, [S, 2S): finished ones, [2S, 3S): compaction chunks arrived (zero between frames)
    float *d_new_T6 = nullptr, *d_st_T6 = nullptr;
    int *d_new_count = nullptr;
    int32_t *d_st_status = nullptr;
    unsigned long long *d_last_id = nullptr, *d_counters = nullptr, *d_rm = nullptr;
    int rm_cap = 0;
    int *d_err = nullptr;  // one overflow flag per slot: a full table is reported for the frame that overflowed only
    unsigned *d_work = nullptr;  // [0,1] temporal LK counters, [2,3] stereo LK counters, [4,5] temporal LK of the new points (split launch)
    int lk_par[3] = {0, 0, 0}, det_par = 0;
    // host-mapped output ring: records [groups][cap] and counts [groups] (+1 error flag)
    OutRecord *h_out[kSlots] = {nullptr, nullptr}, *d_out[kSlots] = {nullptr, nullptr};
    int *h_cnt[kSlots] = {nullptr, nullptr}, *d_cnt[kSlots] = {nullptr, nullptr};
    // unprojection (row f-1): camera models and the host-mapped coordinate ring, allocated when a model is set
    CameraModelDev cam_model[2]{};
    float *h_uxy[kSlots] = {nullptr, nullptr}, *d_uxy[kSlots] = {nullptr, nullptr};
    uint8_t *d_raw = nullptr;  // decoded-frame staging for process_stereo_fmt (2 images), grown on demand
    size_t raw_bytes = 0;
    cudaEvent_t ev_h2d[kSlots]{}, ev_in_free[kSlots]{}, ev_done[kSlots]{}, ev_merged[kSlots]{};
    int next_slot = 0, oldest_slot = 0, inflight = 0, result_slot = -1;
    uint64_t host_launches = 0;
    // profiling
    bool profiling = false;
    cudaEvent_t ev_stage[kSlots][RSVIO_N_STAGES + 1]{};
    cudaEvent_t ev_up[kSlots][2]{};
    bool stage_valid[kSlots] = {false, false}, up_valid[kSlots] = {false, false};
    double stage_ms[RSVIO_N_STAGES] = {0};
    uint64_t stage_launches[RSVIO_N_STAGES] = {0};
    // a CUDA failure may leave a frame half enqueued (tables partly written, events unrecorded): the handle refuses
    // further frames until rsvio_tracker_reset has rebuilt a consistent state
    // CUDA graphs of the per-frame launch sequence (single-stream handles: everything is on one stream there, and the
    // launch overhead of 7 kernels is a third of a frame).  One executable graph per state of the ping-pong buffers.
    struct FrameGraph {
        int slot, pcur, tcur, lk0, lk1, det, have_prev, plan_parity, plan_delta;
        const uint8_t *src0, *src1;
        size_t stride, pitch;
        cudaGraphExec_t exec;
    };
    std::vector<FrameGraph> graphs;
    bool use_graphs = false;
    bool poisoned = false;
    uint64_t overflow_frames = 0;  // frames whose new points did not all fit the track table (sticky count)
    std::string err;
};

namespace {

rsvio_status fail_cuda(rsvio_tracker *t, cudaError_t e, const char *what, int line) {
    cudaGetLastError();  // the error has been reported: do not leave it pending for an unrelated later call
    if (t) {
        char buf[512];
        snprintf(buf, sizeof(buf), "CUDA error %d (%s) at api.cu:%d: %s", (int)e, cudaGetErrorString(e), line, what);
        t->err = buf;
        t->poisoned = true;
    }
    return RSVIO_ERR_CUDA;
}

rsvio_status fail(rsvio_tracker *t, rsvio_status s, const char *msg) {
    if (t) t->err = msg;
    return s;
}

bool is_device_accessible_host(const void *p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return a.type == cudaMemoryTypeHost;
}

template <typename T>
cudaError_t dmalloc(T **p, size_t n) {
    return cudaMalloc((void **)p, n * sizeof(T) > 0 ? n * sizeof(T) : 16);
}

void drop_graphs(rsvio_tracker *t) {
    for (auto &g : t->graphs)
        if (g.exec) cudaGraphExecDestroy(g.exec);
    t->graphs.clear();
}

void free_all(rsvio_tracker *t) {
    if (!t) return;
    DeviceScope dev_scope;
    dev_scope.enter(t->cfg.device);
    if (t->st) cudaStreamSynchronize(t->st);
    if (t->copy_st) cudaStreamSynchronize(t->copy_st);
    if (t->own_exp && t->exp_st) cudaStreamSynchronize(t->exp_st);
    if (t->own_pyr && t->pyr_st) cudaStreamSynchronize(t->pyr_st);
    if (t->lk_st) cudaStreamSynchronize(t->lk_st);
    drop_graphs(t);
    pyr_plan_destroy(t->plan);
    for (int i = 0; i < kSlots; ++i) {
        cudaFree(t->d_in[i]);
        if (t->h_stage[i]) cudaFreeHost(t->h_stage[i]);
        if (t->h_out[i]) cudaFreeHost(t->h_out[i]);
        if (t->h_cnt[i]) cudaFreeHost(t->h_cnt[i]);
        if (t->h_uxy[i]) cudaFreeHost(t->h_uxy[i]);
        if (t->ev_h2d[i]) cudaEventDestroy(t->ev_h2d[i]);
        if (t->ev_in_free[i]) cudaEventDestroy(t->ev_in_free[i]);
        if (t->ev_done[i]) cudaEventDestroy(t->ev_done[i]);
        if (t->ev_merged[i]) cudaEventDestroy(t->ev_merged[i]);
        if (t->ev_pyr[i]) cudaEventDestroy(t->ev_pyr[i]);
        for (int k = 0; k <= RSVIO_N_STAGES; ++k)
            if (t->ev_stage[i][k]) cudaEventDestroy(t->ev_stage[i][k]);
        for (int k = 0; k < 2; ++k)
            if (t->ev_up[i][k]) cudaEventDestroy(t->ev_up[i][k]);
    }
    for (int i = 0; i < 2; ++i) {
        cudaFree(t->d_pyr[i]);
        cudaFree(t->tab[i].ids);
        cudaFree(t->tab[i].T6);
        cudaFree(t->tab[i].counts);
    }
    cudaFree(t->d_raw);
    cudaFree(t->d_lk_T6);
    cudaFree(t->d_lk_status);
    cudaFree(t->d_occ);
    cudaFree(t->d_cell_pt);
    cudaFree(t->d_cell_score);
    cudaFree(t->d_cell_list);
    cudaFree(t->d_cell_list_n);
    cudaFree(t->d_cell_work);
    cudaFree(t->d_new_T6);
    cudaFree(t->d_st_T6);
    cudaFree(t->d_new_count);
    cudaFree(t->d_st_status);
    cudaFree(t->d_last_id);
    cudaFree(t->d_counters);
    cudaFree(t->d_rm);
    cudaFree(t->d_err);
    cudaFree(t->d_work);
    cudaFree(t->d_surv);
    if (t->copy_st) cudaStreamDestroy(t->copy_st);
    if (t->own_exp && t->exp_st) cudaStreamDestroy(t->exp_st);
    if (t->own_pyr && t->pyr_st) cudaStreamDestroy(t->pyr_st);
    if (t->lk_st) cudaStreamDestroy(t->lk_st);
    if (t->ev_compact) cudaEventDestroy(t->ev_compact);
    if (t->ev_lk1) cudaEventDestroy(t->ev_lk1);
    if (t->ev_lk_done) cudaEventDestroy(t->ev_lk_done);
    if (t->own_stream && t->st) cudaStreamDestroy(t->st);
    cudaGetLastError();
    delete t;
}

rsvio_status clear_state(rsvio_tracker *t) {
    for (int i = 0; i < 2; ++i) CU_TRY(t, cudaMemsetAsync(t->tab[i].counts, 0, sizeof(int) * t->groups, t->st));
    CU_TRY(t, cudaMemsetAsync(t->d_last_id, 0, sizeof(unsigned long long) * t->S, t->st));
    CU_TRY(t, cudaMemsetAsync(t->d_counters, 0, sizeof(unsigned long long) * CNT_N, t->st));
    CU_TRY(t, cudaMemsetAsync(t->d_err, 0, sizeof(int) * kSlots, t->st));
    CU_TRY(t, cudaMemsetAsync(t->d_new_count, 0, sizeof(int) * t->S, t->st));
    CU_TRY(t, cudaMemsetAsync(t->d_work, 0, sizeof(unsigned) * 6, t->st));
    CU_TRY(t, cudaMemsetAsync(t->d_surv, 0, sizeof(int) * t->groups, t->st));
    CU_TRY(t, cudaMemsetAsync(t->d_cell_list_n, 0, sizeof(int) * 2, t->st));
    CU_TRY(t, cudaMemsetAsync(t->d_cell_work, 0, sizeof(int) * 3 * t->S, t->st));
    // the occupancy grids start clear; every compaction launch leaves them clear again (tracks.cu)
    CU_TRY(t, cudaMemsetAsync(t->d_occ, 0, (size_t)t->S * t->gd.occ_rows * t->gd.occ_cols, t->st));
    t->lk_par[0] = t->lk_par[1] = t->lk_par[2] = t->det_par = 0;
    t->surv_valid = false;
    t->upload_bound = false;
    if (t->lk_st) CU_TRY(t, cudaStreamSynchronize(t->lk_st));
    if (t->own_exp) CU_TRY(t, cudaStreamSynchronize(t->exp_st));
    if (t->own_pyr) CU_TRY(t, cudaStreamSynchronize(t->pyr_st));
    CU_TRY(t, cudaStreamSynchronize(t->st));
    for (int i = 0; i < kSlots; ++i) memset(t->h_cnt[i], 0, sizeof(int) * (t->groups + 1));
    t->have_prev = false;
    t->tcur = 0;
    t->pcur = 0;
    t->inflight = 0;
    t->next_slot = t->oldest_slot = 0;
    t->result_slot = -1;
    t->host_launches = 0;
    t->overflow_frames = 0;
    t->poisoned = false;
    return RSVIO_OK;
}

// The launches of one frame on t->st (and the side streams of batched handles).  src0/src1: device images of camera
// 0 / 1 (S each).  Changes no host state of the handle: frame_advance() does, once the launches are out.  With
// `capturing` the stream is being captured into a graph: everything is on t->st and no event is recorded (the
// caller records them after the graph launch).
rsvio_status frame_launches(rsvio_tracker *t, int slot, const uint8_t *src0, const uint8_t *src1, size_t stride,
                            size_t pitch, bool capturing) {
    cudaStream_t st = t->st;
    const bool prof = t->profiling && !capturing;
    int stage = 1;  // stage 0 (upload) is timed on the copy stream
    auto mark = [&](int k) {
        if (prof) cudaEventRecord(t->ev_stage[slot][k], st);
    };
    auto count = [&](int stg, int n) {
        if (capturing) return;  // replays are counted by the caller
        t->host_launches += (uint64_t)n;
        t->stage_launches[stg] += (uint64_t)n;
    };
    uint8_t *cur = t->d_pyr[t->pcur], *prev = t->d_pyr[t->pcur ^ 1];
    const int n_img = t->stereo ? 2 * t->S : t->S;
    const TableView &tin = t->tab[t->tcur], &tout = t->tab[t->tcur ^ 1];

    // 1: pyramid.  On batched handles it runs on its own stream, so that the pyramid of this frame overlaps
    // the low-occupancy tail (compaction, detection, stereo LK, merge) of the previous one.  It overwrites the
    // set that was the previous frame's `prev`, last read by that frame's temporal LK.  Profiling keeps
    // everything on one stream (stage times would overlap otherwise).
    cudaStream_t ps = (prof || capturing) ? st : t->pyr_st;
    if (ps != st) CU_TRY(t, cudaStreamWaitEvent(ps, t->ev_lk_done, 0));
    mark(stage);
    int nl = 0;
    CU_TRY(t, pyr_build(t->plan, src0, src1, t->S, n_img, stride, pitch, cur, ps, &nl));
    count(1, nl);
    if (!capturing) CU_TRY(t, cudaEventRecord(t->ev_in_free[slot], ps));
    if (ps != st) {
        CU_TRY(t, cudaEventRecord(t->ev_pyr[slot], ps));
        CU_TRY(t, cudaStreamWaitEvent(st, t->ev_pyr[slot], 0));
    }
    mark(++stage);  // 2: temporal LK
    // split launch (see lk_st): rows [0, surv) on lk_st behind the previous compaction, rows [surv, count) here
    const bool split = t->split_lk && !prof && !capturing && t->have_prev && t->surv_valid && !t->upload_bound;
    t->last_split = split;
    if (t->have_prev) {
        LkBatch b{};
        b.pyr_a = prev;
        b.pyr_b = cur;
        b.counts = tin.counts;
        b.in_T6 = tin.T6;
        b.out_T6 = t->d_lk_T6;
        b.status = t->d_lk_status;
        b.iters = nullptr;
        b.counters = t->d_counters;
        b.work_counter = t->d_work;
        b.work_parity = t->lk_par[0];
        b.groups = t->stereo ? t->groups : t->S;
        b.cap = t->cap;
        b.max_it = t->cfg.max_iterations;
        b.thr = t->cfg.convergence_threshold;
        if (split) {
            // the bulk: one CTA slot per SM stays free for the previous frame's detection, stereo LK, merge and export
            LkBatch b1 = b;
            b1.counts = t->d_surv;
            b1.blocks_per_sm = 7;
            CU_TRY(t, cudaStreamWaitEvent(t->lk_st, t->ev_pyr[slot], 0));
            CU_TRY(t, cudaStreamWaitEvent(t->lk_st, t->ev_compact, 0));
            CU_TRY(t, lk_track(t->pd, b1, t->lk_st));
            CU_TRY(t, cudaEventRecord(t->ev_lk1, t->lk_st));
            // the points the previous frame's merge appended (st order: behind that merge)
            b.first = t->d_surv;
            b.work_counter = t->d_work + 4;
            b.work_parity = t->lk_par[2];
            CU_TRY(t, lk_track(t->pd, b, st));
            // By now (st order) the previous frame's tail is done and the free CTA slot would idle until the bulk launch
            // ends: one more persistent CTA per SM joins it, drawing from the same work counter (RSVIO_LK_HELPER=0: off)
            if (t->lk_helper) {
                b1.blocks_per_sm = 1;
                CU_TRY(t, lk_track(t->pd, b1, st));
            }
            CU_TRY(t, cudaStreamWaitEvent(st, t->ev_lk1, 0));
            count(2, t->lk_helper ? 3 : 2);
        } else {
            CU_TRY(t, lk_track(t->pd, b, st));
            count(2, 1);
        }
    }
    if (!capturing) CU_TRY(t, cudaEventRecord(t->ev_lk_done, st));  // `prev` may be overwritten from here on
    mark(++stage);  // 3: compaction + occupancy + the list of empty cells
    CellWork cw{};
    cw.cell_pt = t->d_cell_pt;
    cw.cell_score = t->d_cell_score;
    cw.list = t->d_cell_list;
    cw.list_count = t->d_cell_list_n + t->det_par;
    cw.next_count = t->d_cell_list_n + (t->det_par ^ 1);
    cw.n_empty = t->d_cell_work;
    cw.done = t->d_cell_work + t->S;
    cw.arrive = t->d_cell_work + 2 * t->S;
    cw.new_count = t->d_new_count;
    CU_TRY(t, compact_tracks(tin, t->d_lk_T6, t->d_lk_status, tout, t->groups, t->cap, t->S, t->gd, t->d_occ,
                             t->have_prev ? 1 : 0, t->d_counters, t->d_err + slot, cw, t->d_surv, st));
    if (t->split_lk && !capturing) CU_TRY(t, cudaEventRecord(t->ev_compact, st));
    count(3, 1);
    mark(++stage);  // 4: detection (level 0 of the current left pyramid is the input image) + ordered collection
    CU_TRY(t, detect_cells(t->gd, cur + t->pd.lv[0].off, (size_t)t->pd.lv[0].w, (size_t)t->pd.pyr_stride, t->S,
                           t->d_occ, cw, t->d_new_T6, nullptr, t->new_cap, t->d_counters, false, t->split_lk && !prof && !t->upload_bound, st));
    count(4, 1);
    mark(++stage);  // 5: stereo LK (left -> right of the current frame)
    if (t->stereo) {
        LkBatch b{};
        b.pyr_a = cur;
        b.pyr_b = cur + (size_t)t->S * t->pd.pyr_stride;
        b.counts = t->d_new_count;
        b.in_T6 = t->d_new_T6;
        b.out_T6 = t->d_st_T6;
        b.status = t->d_st_status;
        b.iters = nullptr;
        b.counters = t->d_counters;
        b.work_counter = t->d_work + 2;
        b.work_parity = t->lk_par[1];
        b.groups = t->S;
        b.cap = t->new_cap;
        b.max_it = t->cfg.max_iterations;
        b.thr = t->cfg.convergence_threshold;
        CU_TRY(t, lk_track(t->pd, b, st));
        count(5, 1);
    }
    mark(++stage);  // 6: merge (ids, table append)
    CU_TRY(t, merge_new_points(tout, t->S, t->cap, t->stereo, t->d_new_T6, t->d_st_T6, t->d_st_status,
                               t->d_new_count, t->new_cap, t->d_last_id, t->d_err + slot, st));
    count(6, 1);
    // 7: download -- packed records into mapped host memory + the error flag.  With many streams this runs on
    // its own CUDA stream beside the next frame (the table it reads is not written again before the frame after
    // next, which cannot be submitted until this one has been synced).
    cudaStream_t xs = (prof || capturing) ? st : t->exp_st;  // profiling: one stream, every stage alone on the device
    if (xs != st) {
        CU_TRY(t, cudaEventRecord(t->ev_merged[slot], st));
        CU_TRY(t, cudaStreamWaitEvent(xs, t->ev_merged[slot], 0));
    }
    ++stage;
    if (prof) cudaEventRecord(t->ev_stage[slot][stage], xs);
    CU_TRY(t, export_records(tout, t->groups, t->cap, t->S, t->d_out[slot], t->d_cnt[slot], t->cam_model[0], t->cam_model[1],
                             t->d_uxy[slot], t->d_err + slot, xs));
    count(7, 1);
    ++stage;
    if (prof) cudaEventRecord(t->ev_stage[slot][stage], xs);  // end
    if (!capturing) {
        t->stage_valid[slot] = prof;
        CU_TRY(t, cudaEventRecord(t->ev_done[slot], xs));
    }
    return RSVIO_OK;
}

// host state after the launches of a frame went out: the ping-pong sets and the counter pairs switch sides
void frame_advance(rsvio_tracker *t) {
    if (t->have_prev) t->lk_par[0] ^= 1;  // only a launch that went out has cleared the other counter
    if (t->last_split) t->lk_par[2] ^= 1;
    t->surv_valid = true;  // the compaction of this frame wrote d_surv for the table the next frame tracks
    if (t->stereo) t->lk_par[1] ^= 1;
    t->det_par ^= 1;
    t->tcur ^= 1;
    t->pcur ^= 1;
    t->have_prev = true;
}

// Single-stream handles: the frame as one CUDA graph launch.  Graphs are captured lazily, one per state of the
// ping-pong buffers (two in steady state), and replayed while that state recurs.
rsvio_status enqueue_frame_graph(rsvio_tracker *t, int slot, const uint8_t *src0, const uint8_t *src1, size_t stride,
                                 size_t pitch) {
    const int parity = pyr_plan_parity(t->plan);
    rsvio_tracker::FrameGraph *hit = nullptr;
    for (auto &g : t->graphs)
        if (g.slot == slot && g.pcur == t->pcur && g.tcur == t->tcur && g.lk0 == t->lk_par[0] && g.lk1 == t->lk_par[1] &&
            g.det == t->det_par && g.have_prev == (t->have_prev ? 1 : 0) && g.plan_parity == parity && g.src0 == src0 &&
            g.src1 == src1 && g.stride == stride && g.pitch == pitch) {
            hit = &g;
            break;
        }
    if (!hit) {
        if (t->graphs.size() >= 16) drop_graphs(t);  // callers that keep changing their buffers: start over
        CU_TRY(t, cudaStreamBeginCapture(t->st, cudaStreamCaptureModeThreadLocal));
        const rsvio_status rs = frame_launches(t, slot, src0, src1, stride, pitch, true);
        cudaGraph_t graph = nullptr;
        const cudaError_t ec = cudaStreamEndCapture(t->st, &graph);
        const int after = pyr_plan_parity(t->plan);
        pyr_plan_set_parity(t->plan, parity);  // nothing ran yet: the host copy follows the replays below
        if (rs != RSVIO_OK || ec != cudaSuccess || !graph) {
            if (graph) cudaGraphDestroy(graph);
            return rs != RSVIO_OK ? rs : fail_cuda(t, ec != cudaSuccess ? ec : cudaErrorUnknown, "cudaStreamEndCapture", __LINE__);
        }
        rsvio_tracker::FrameGraph g{slot, t->pcur, t->tcur, t->lk_par[0], t->lk_par[1], t->det_par, t->have_prev ? 1 : 0,
                                    parity, after ^ parity, src0, src1, stride, pitch, nullptr};
        const cudaError_t ei = cudaGraphInstantiate(&g.exec, graph, 0);
        cudaGraphDestroy(graph);
        if (ei != cudaSuccess) return fail_cuda(t, ei, "cudaGraphInstantiate", __LINE__);
        t->graphs.push_back(g);
        hit = &t->graphs.back();
    }
    CU_TRY(t, cudaGraphLaunch(hit->exec, t->st));
    pyr_plan_set_parity(t->plan, parity ^ hit->plan_delta);
    const int n_launch = 5 + (t->have_prev ? 1 : 0) + (t->stereo ? 1 : 0);
    t->host_launches += (uint64_t)n_launch;
    t->stage_launches[1] += 1;
    CU_TRY(t, cudaEventRecord(t->ev_in_free[slot], t->st));
    CU_TRY(t, cudaEventRecord(t->ev_lk_done, t->st));
    t->stage_valid[slot] = false;
    CU_TRY(t, cudaEventRecord(t->ev_done[slot], t->st));
    return RSVIO_OK;
}

rsvio_status enqueue_frame(rsvio_tracker *t, int slot, const uint8_t *src0, const uint8_t *src1, size_t stride,
                           size_t pitch) {
    rsvio_status rs;
    if (t->use_graphs && !t->profiling && t->pyr_st == t->st && t->exp_st == t->st)
        rs = enqueue_frame_graph(t, slot, src0, src1, stride, pitch);
    else
        rs = frame_launches(t, slot, src0, src1, stride, pitch, false);
    if (rs != RSVIO_OK) return rs;
    frame_advance(t);
    return RSVIO_OK;
}

rsvio_status upload(rsvio_tracker *t, int slot, const uint8_t *const *left, const uint8_t *const *right,
                    size_t stride) {
    const int w = t->cfg.width, h = t->cfg.height, S = t->S;
    const int ncam = t->stereo ? 2 : 1;
    // the copy engine must not overwrite d_in[slot] while the pyramid kernel of the frame that
    // last used this slot may still read it
    CU_TRY(t, cudaStreamWaitEvent(t->copy_st, t->ev_in_free[slot], 0));
    if (t->profiling) cudaEventRecord(t->ev_up[slot][0], t->copy_st);
    for (int cam = 0; cam < ncam; ++cam) {
        const uint8_t *const *src = cam == 0 ? left : right;
        uint8_t *dst = t->d_in[slot] + (size_t)cam * S * t->img_bytes;
        const bool pinned = is_device_accessible_host(src[0]);
        // contiguous batch: constant pitch between consecutive images
        bool uniform = true;
        const ptrdiff_t pitch = S > 1 ? src[1] - src[0] : (ptrdiff_t)(stride * h);
        for (int s = 1; s < S && uniform; ++s) uniform = (src[s] - src[s - 1]) == pitch;
        if (!pinned) {
            // pageable source: gather into the slot's pinned staging buffer (the caller may free
            // or overwrite its images as soon as this call returns), then one DMA.
            CU_TRY(t, cudaEventSynchronize(t->ev_h2d[slot]));
            if (!t->h_stage[slot])  // allocated on first use: callers with pinned images never need it
                CU_TRY(t, cudaHostAlloc((void **)&t->h_stage[slot], 2 * (size_t)S * t->img_bytes, cudaHostAllocDefault));
            uint8_t *stage = t->h_stage[slot] + (size_t)cam * S * t->img_bytes;
            for (int s = 0; s < S; ++s)
                for (int y = 0; y < h; ++y)
                    memcpy(stage + (size_t)s * t->img_bytes + (size_t)y * w, src[s] + (size_t)y * stride, (size_t)w);
            CU_TRY(t, cudaMemcpyAsync(dst, stage, (size_t)S * t->img_bytes, cudaMemcpyHostToDevice, t->copy_st));
            t->stage_launches[0] += 1;
        } else if (uniform && stride == (size_t)w && pitch == (ptrdiff_t)t->img_bytes) {
            CU_TRY(t, cudaMemcpyAsync(dst, src[0], (size_t)S * t->img_bytes, cudaMemcpyHostToDevice, t->copy_st));
            t->stage_launches[0] += 1;
        } else {
            for (int s = 0; s < S; ++s)
                CU_TRY(t, cudaMemcpy2DAsync(dst + (size_t)s * t->img_bytes, (size_t)w, src[s], stride, (size_t)w,
                                            (size_t)h, cudaMemcpyHostToDevice, t->copy_st));
            t->stage_launches[0] += (uint64_t)S;
        }
    }
    if (t->profiling) cudaEventRecord(t->ev_up[slot][1], t->copy_st);
    t->up_valid[slot] = t->profiling;
    CU_TRY(t, cudaEventRecord(t->ev_h2d[slot], t->copy_st));
    t->slot_host[slot] = true;
    CU_TRY(t, cudaStreamWaitEvent(t->st, t->ev_h2d[slot], 0));
    if (t->pyr_st != t->st) CU_TRY(t, cudaStreamWaitEvent(t->pyr_st, t->ev_h2d[slot], 0));
    return RSVIO_OK;
}

rsvio_status check_batch_args(rsvio_tracker *t, const uint8_t *const *left, const uint8_t *const *right,
                              size_t stride) {
    if (!t) return RSVIO_ERR_INVALID_ARG;
    if (!left || (t->stereo && !right)) return fail(t, RSVIO_ERR_INVALID_ARG, "null image array");
    if (stride < (size_t)t->cfg.width) return fail(t, RSVIO_ERR_INVALID_ARG, "stride smaller than width");
    for (int s = 0; s < t->S; ++s)
        if (!left[s] || (t->stereo && !right[s])) return fail(t, RSVIO_ERR_INVALID_ARG, "null image pointer");
    if (t->poisoned) return fail(t, RSVIO_ERR_STATE, "an earlier CUDA error left the handle inconsistent: call rsvio_tracker_reset");
    if (t->inflight >= kSlots) return fail(t, RSVIO_ERR_STATE, "two frames already in flight: call rsvio_tracker_sync");
    return RSVIO_OK;
}

}  // namespace

extern "C" {

int32_t rsvio_abi_version(void) { return RSVIO_ABI_VERSION; }

const char *rsvio_status_string(rsvio_status s) {
    switch (s) {
        case RSVIO_OK: return "ok";
        case RSVIO_ERR_INVALID_ARG: return "invalid argument";
        case RSVIO_ERR_CUDA: return "CUDA error or no usable device";
        case RSVIO_ERR_CAPACITY: return "capacity exceeded";
        case RSVIO_ERR_STATE: return "invalid call order";
        case RSVIO_ERR_ALLOC: return "allocation failed";
        default: return "unknown status";
    }
}

rsvio_status rsvio_device_count(int32_t *n) {
    if (!n) return RSVIO_ERR_INVALID_ARG;
    int c = 0;
    if (cudaGetDeviceCount(&c) != cudaSuccess) {
        cudaGetLastError();
        *n = 0;
        return RSVIO_ERR_CUDA;
    }
    *n = c;
    return c > 0 ? RSVIO_OK : RSVIO_ERR_CUDA;
}

void rsvio_config_default(rsvio_config *cfg) {
    if (!cfg) return;
    memset(cfg, 0, sizeof(*cfg));
    cfg->levels = 6;                       /* estimator.rs:27 */
    cfg->grid_size = 30;                   /* feature_tracker.rs:65 */
    cfg->max_iterations = 30;              /* feature_tracker.rs:51 */
    cfg->convergence_threshold = 0.005f;   /* feature_tracker.rs:52 */
    cfg->mode = RSVIO_MODE_STEREO;
    cfg->n_streams = 1;
}

rsvio_status rsvio_host_alloc(void **ptr, size_t bytes) {
    if (!ptr) return RSVIO_ERR_INVALID_ARG;
    if (cudaHostAlloc(ptr, bytes ? bytes : 1, cudaHostAllocPortable) != cudaSuccess) {
        cudaGetLastError();
        *ptr = nullptr;
        return RSVIO_ERR_CUDA;
    }
    return RSVIO_OK;
}

rsvio_status rsvio_host_alloc_flags(void **ptr, size_t bytes, uint32_t flags) {
    if (!ptr || (flags & ~(uint32_t)RSVIO_HOST_WRITE_COMBINED)) return RSVIO_ERR_INVALID_ARG;
    unsigned f = cudaHostAllocPortable;
    if (flags & RSVIO_HOST_WRITE_COMBINED) f |= cudaHostAllocWriteCombined;
    if (cudaHostAlloc(ptr, bytes ? bytes : 1, f) != cudaSuccess) {
        cudaGetLastError();
        *ptr = nullptr;
        return RSVIO_ERR_CUDA;
    }
    return RSVIO_OK;
}

rsvio_status rsvio_host_free(void *ptr) {
    if (!ptr) return RSVIO_OK;
    if (cudaFreeHost(ptr) != cudaSuccess) {
        cudaGetLastError();
        return RSVIO_ERR_CUDA;
    }
    return RSVIO_OK;
}

rsvio_status rsvio_tracker_create(const rsvio_config *cfg, rsvio_tracker **out) {
    if (!cfg || !out) return RSVIO_ERR_INVALID_ARG;
    *out = nullptr;
    if (cfg->width < 8 || cfg->height < 8 || cfg->width > 65535 || cfg->height > 65535 || cfg->levels < 1 ||
        cfg->levels > RSVIO_MAX_LEVELS || (cfg->width >> (cfg->levels - 1)) < 1 ||
        (cfg->height >> (cfg->levels - 1)) < 1 || cfg->grid_size < 7 || cfg->grid_size > 1024 ||
        cfg->max_iterations < 0 || cfg->n_streams < 1 || cfg->max_tracks < 0 ||
        !std::isfinite(cfg->convergence_threshold) ||  // NaN never compares; a negative one is legal (never converges early)
        (cfg->mode != RSVIO_MODE_STEREO && cfg->mode != RSVIO_MODE_MONO))
        return RSVIO_ERR_INVALID_ARG;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || cfg->device < 0 || cfg->device >= ndev) {
        cudaGetLastError();
        return RSVIO_ERR_CUDA;
    }
    DeviceScope dev_scope;
    if (dev_scope.enter(cfg->device) != cudaSuccess) return RSVIO_ERR_CUDA;
    rsvio_tracker *t = new (std::nothrow) rsvio_tracker();
    if (!t) return RSVIO_ERR_ALLOC;
    t->cfg = *cfg;
    t->S = cfg->n_streams;
    t->groups = 2 * t->S;
    t->stereo = cfg->mode == RSVIO_MODE_STEREO ? 1 : 0;
    t->gd = make_grid_desc(cfg->width, cfg->height, cfg->grid_size);
    const int cells = t->gd.ncx * t->gd.ncy;
    t->new_cap = cells > 0 ? cells : 1;
    t->cap = cfg->max_tracks > 0 ? cfg->max_tracks : 4 * cells + 64;
    t->img_bytes = (size_t)cfg->width * cfg->height;
    rsvio_status rs = RSVIO_OK;
    auto body = [&]() -> rsvio_status {
        if (cfg->cuda_stream) {
            t->st = (cudaStream_t)cfg->cuda_stream;
        } else {
            CU_TRY(t, cudaStreamCreateWithFlags(&t->st, cudaStreamNonBlocking));
            t->own_stream = true;
        }
        CU_TRY(t, cudaStreamCreateWithFlags(&t->copy_st, cudaStreamNonBlocking));
        if (t->S >= 8) {  // small batches: the cross-stream hand-offs would cost more latency than they hide
            CU_TRY(t, cudaStreamCreateWithFlags(&t->exp_st, cudaStreamNonBlocking));
            t->own_exp = true;
            {   // lowest priority: the pyramid fills what the latency-bound tail of the previous frame leaves idle
                int lo = 0, hi = 0;
                CU_TRY(t, cudaDeviceGetStreamPriorityRange(&lo, &hi));
                CU_TRY(t, cudaStreamCreateWithPriority(&t->pyr_st, cudaStreamNonBlocking, lo));
            }
            t->own_pyr = true;
            // the split pays when the bulk launch outlasts the previous frame's latency-bound tail (~80 us, ~5.5 us
            // of temporal LK per EuRoC-size stream): from about 24 streams up
            const char *ns = getenv("RSVIO_NO_SPLIT_LK");
            if (t->S >= 24 && !(ns && ns[0] == '1')) {
                CU_TRY(t, cudaStreamCreateWithFlags(&t->lk_st, cudaStreamNonBlocking));
                CU_TRY(t, cudaEventCreateWithFlags(&t->ev_compact, cudaEventDisableTiming));
                CU_TRY(t, cudaEventCreateWithFlags(&t->ev_lk1, cudaEventDisableTiming));
                t->split_lk = true;
                const char *hp = getenv("RSVIO_LK_HELPER");
                t->lk_helper = !(hp && hp[0] == '0');
            }
        } else {
            t->exp_st = t->st;
            t->pyr_st = t->st;
            const char *ng = getenv("RSVIO_NO_GRAPH");
            t->use_graphs = !(ng && ng[0] == '1');
        }
        CU_TRY(t, cudaEventCreateWithFlags(&t->ev_lk_done, cudaEventDisableTiming));
        CU_TRY(t, pyr_plan_create(cfg->width, cfg->height, cfg->levels, t->st, &t->plan));
        CU_TRY(t, pyr_plan_reserve(t->plan, 2 * t->S));
        t->pd = pyr_plan_desc(t->plan);
        const size_t G = (size_t)t->groups, cap = (size_t)t->cap, S = (size_t)t->S;
        for (int i = 0; i < kSlots; ++i) {
            CU_TRY(t, dmalloc(&t->d_in[i], 2 * S * t->img_bytes));
            CU_TRY(t, cudaHostAlloc((void **)&t->h_out[i], G * cap * sizeof(OutRecord), cudaHostAllocMapped));
            CU_TRY(t, cudaHostGetDevicePointer((void **)&t->d_out[i], t->h_out[i], 0));
            CU_TRY(t, cudaHostAlloc((void **)&t->h_cnt[i], (G + 1) * sizeof(int), cudaHostAllocMapped));
            CU_TRY(t, cudaHostGetDevicePointer((void **)&t->d_cnt[i], t->h_cnt[i], 0));
            CU_TRY(t, cudaEventCreateWithFlags(&t->ev_h2d[i], cudaEventDisableTiming));
            CU_TRY(t, cudaEventCreateWithFlags(&t->ev_in_free[i], cudaEventDisableTiming));
            CU_TRY(t, cudaEventCreateWithFlags(&t->ev_done[i], cudaEventDisableTiming));
            CU_TRY(t, cudaEventCreateWithFlags(&t->ev_merged[i], cudaEventDisableTiming));
            CU_TRY(t, cudaEventCreateWithFlags(&t->ev_pyr[i], cudaEventDisableTiming));
            for (int k = 0; k <= RSVIO_N_STAGES; ++k) CU_TRY(t, cudaEventCreate(&t->ev_stage[i][k]));
            for (int k = 0; k < 2; ++k) CU_TRY(t, cudaEventCreate(&t->ev_up[i][k]));
        }
        for (int i = 0; i < 2; ++i) {
            CU_TRY(t, dmalloc(&t->d_pyr[i], G * (size_t)t->pd.pyr_stride));
            CU_TRY(t, dmalloc(&t->tab[i].ids, G * cap));
            CU_TRY(t, dmalloc(&t->tab[i].T6, G * cap * 6));
            CU_TRY(t, dmalloc(&t->tab[i].counts, G));
        }
        CU_TRY(t, dmalloc(&t->d_lk_T6, G * cap * 6));
        CU_TRY(t, dmalloc(&t->d_lk_status, G * cap));
        CU_TRY(t, dmalloc(&t->d_occ, S * (size_t)t->gd.occ_rows * t->gd.occ_cols));
        CU_TRY(t, dmalloc(&t->d_cell_pt, S * (size_t)t->new_cap));
        CU_TRY(t, dmalloc(&t->d_cell_score, S * (size_t)t->new_cap));
        CU_TRY(t, dmalloc(&t->d_cell_list, S * (size_t)t->new_cap));
        CU_TRY(t, dmalloc(&t->d_cell_list_n, (size_t)2));
        CU_TRY(t, dmalloc(&t->d_cell_work, 3 * S));
        CU_TRY(t, dmalloc(&t->d_new_T6, S * (size_t)t->new_cap * 6));
        CU_TRY(t, dmalloc(&t->d_st_T6, S * (size_t)t->new_cap * 6));
        CU_TRY(t, dmalloc(&t->d_new_count, S));
        CU_TRY(t, dmalloc(&t->d_st_status, S * (size_t)t->new_cap));
        CU_TRY(t, dmalloc(&t->d_last_id, S));
        CU_TRY(t, dmalloc(&t->d_counters, (size_t)CNT_N));
        CU_TRY(t, dmalloc(&t->d_err, (size_t)kSlots));
        CU_TRY(t, dmalloc(&t->d_work, (size_t)6));
        CU_TRY(t, dmalloc(&t->d_surv, G));
        return clear_state(t);
    };
    rs = body();
    if (rs != RSVIO_OK) {
        fprintf(stderr, "rsvio_tracker_create failed: %s\n", t->err.c_str());
        free_all(t);
        return rs;
    }
    *out = t;
    return RSVIO_OK;
}

void rsvio_tracker_destroy(rsvio_tracker *t) { free_all(t); }

const char *rsvio_tracker_last_error(const rsvio_tracker *t) { return t ? t->err.c_str() : "null handle"; }

rsvio_status rsvio_tracker_reset(rsvio_tracker *t) {
    if (!t) return RSVIO_ERR_INVALID_ARG;
    DeviceScope dev_scope;
    CU_TRY(t, dev_scope.enter(t->cfg.device));
    CU_TRY(t, cudaStreamSynchronize(t->copy_st));
    CU_TRY(t, cudaStreamSynchronize(t->st));
    return clear_state(t);
}

rsvio_status rsvio_tracker_submit_batch(rsvio_tracker *t, const uint8_t *const *left, const uint8_t *const *right,
                                        size_t stride) {
    rsvio_status rs = check_batch_args(t, left, right, stride);
    if (rs != RSVIO_OK) return rs;
    DeviceScope dev_scope;
    CU_TRY(t, dev_scope.enter(t->cfg.device));
    const int slot = t->next_slot;
    rs = upload(t, slot, left, right, stride);
    if (rs != RSVIO_OK) return rs;
    const uint8_t *d0 = t->d_in[slot], *d1 = t->d_in[slot] + (size_t)t->S * t->img_bytes;
    rs = enqueue_frame(t, slot, d0, d1, (size_t)t->cfg.width, t->img_bytes);
    if (rs != RSVIO_OK) return rs;
    t->next_slot ^= 1;
    t->inflight += 1;
    return RSVIO_OK;
}

rsvio_status rsvio_tracker_submit_device(rsvio_tracker *t, const void *d_left, const void *d_right, size_t stride,
                                         size_t pitch) {
    if (!t) return RSVIO_ERR_INVALID_ARG;
    if (!d_left || (t->stereo && !d_right)) return fail(t, RSVIO_ERR_INVALID_ARG, "null device image pointer");
    if (stride < (size_t)t->cfg.width || pitch < stride * (size_t)(t->cfg.height - 1) + (size_t)t->cfg.width)
        return fail(t, RSVIO_ERR_INVALID_ARG, "stride/pitch too small");
    if (t->poisoned) return fail(t, RSVIO_ERR_STATE, "an earlier CUDA error left the handle inconsistent: call rsvio_tracker_reset");
    if (t->inflight >= kSlots) return fail(t, RSVIO_ERR_STATE, "two frames already in flight: call rsvio_tracker_sync");
    DeviceScope dev_scope;
    CU_TRY(t, dev_scope.enter(t->cfg.device));
    const int slot = t->next_slot;
    t->up_valid[slot] = false;
    t->slot_host[slot] = false;
    rsvio_status rs = enqueue_frame(t, slot, (const uint8_t *)d_left, (const uint8_t *)d_right, stride, pitch);
    if (rs != RSVIO_OK) return rs;
    t->next_slot ^= 1;
    t->inflight += 1;
    return RSVIO_OK;
}

rsvio_status rsvio_tracker_sync(rsvio_tracker *t) {
    if (!t) return RSVIO_ERR_INVALID_ARG;
    if (t->inflight <= 0) return fail(t, RSVIO_ERR_STATE, "nothing in flight");
    DeviceScope dev_scope;
    CU_TRY(t, dev_scope.enter(t->cfg.device));
    const int slot = t->oldest_slot;
    {
        const cudaError_t e = cudaEventSynchronize(t->ev_done[slot]);
        if (e != cudaSuccess) {  // the frame is lost: retire it (so a reset is possible) and poison the handle
            t->oldest_slot ^= 1;
            t->inflight -= 1;
            return fail_cuda(t, e, "cudaEventSynchronize(ev_done)", __LINE__);
        }
    }
    if (t->split_lk && t->inflight >= 2) {  // a frame has been submitted behind this one: was the device waiting for it?
        const int nslot = slot ^ 1;
        t->upload_bound = t->slot_host[nslot] && cudaEventQuery(t->ev_h2d[nslot]) == cudaErrorNotReady;
        cudaGetLastError();
    }
    if (t->stage_valid[slot]) {
        for (int k = 1; k < RSVIO_N_STAGES; ++k) {
            float ms = 0.0f;
            if (cudaEventElapsedTime(&ms, t->ev_stage[slot][k], t->ev_stage[slot][k + 1]) == cudaSuccess)
                t->stage_ms[k] += ms;
        }
        t->stage_valid[slot] = false;
    }
    if (t->up_valid[slot]) {
        float ms = 0.0f;
        if (cudaEventElapsedTime(&ms, t->ev_up[slot][0], t->ev_up[slot][1]) == cudaSuccess) t->stage_ms[0] += ms;
        t->up_valid[slot] = false;
    }
    cudaGetLastError();
    t->oldest_slot ^= 1;
    t->inflight -= 1;
    t->result_slot = slot;
    if (t->h_cnt[slot][t->groups] != 0) {
        // not fatal: the frame has been processed and its result is valid, only the new points that did not fit
        // were dropped (the reference's maps are unbounded).  The status tells the caller for THIS frame; the count
        // stays queryable (rsvio_tracker_overflow_frames).
        t->overflow_frames += 1;
        return fail(t, RSVIO_ERR_CAPACITY, "track table full, new points dropped: raise rsvio_config.max_tracks");
    }
    return RSVIO_OK;
}

rsvio_status rsvio_tracker_process_batch(rsvio_tracker *t, const uint8_t *const *left, const uint8_t *const *right,
                                         size_t stride) {
    rsvio_status rs = rsvio_tracker_submit_batch(t, left, right, stride);
    if (rs != RSVIO_OK) return rs;
    return rsvio_tracker_sync(t);
}

rsvio_status rsvio_tracker_process_stereo(rsvio_tracker *t, const uint8_t *left, const uint8_t *right, size_t stride) {
    if (!t) return RSVIO_ERR_INVALID_ARG;
    if (t->S != 1 || !t->stereo) return fail(t, RSVIO_ERR_INVALID_ARG, "handle is not a 1-stream stereo tracker");
    const uint8_t *l[1] = {left}, *r[1] = {right};
    return rsvio_tracker_process_batch(t, l, r, stride);
}

rsvio_status rsvio_tracker_process_stereo_fmt(rsvio_tracker *t, const void *left, const void *right,
                                              size_t stride_bytes, int32_t pixel_format) {
    if (!t) return RSVIO_ERR_INVALID_ARG;
    if (pixel_format == RSVIO_PIXEL_LUMA8)
        return rsvio_tracker_process_stereo(t, (const uint8_t *)left, (const uint8_t *)right, stride_bytes);
    if (t->S != 1 || !t->stereo) return fail(t, RSVIO_ERR_INVALID_ARG, "handle is not a 1-stream stereo tracker");
    const int bpp = pixel_bytes(pixel_format);
    if (bpp == 0) return fail(t, RSVIO_ERR_INVALID_ARG, "unknown pixel format");
    if (!left || !right) return fail(t, RSVIO_ERR_INVALID_ARG, "null image pointer");
    const int w = t->cfg.width, h = t->cfg.height;
    const size_t row = (size_t)w * bpp;
    if (stride_bytes < row) return fail(t, RSVIO_ERR_INVALID_ARG, "stride smaller than a row");
    if (t->poisoned) return fail(t, RSVIO_ERR_STATE, "an earlier CUDA error left the handle inconsistent: call rsvio_tracker_reset");
    if (t->inflight >= kSlots) return fail(t, RSVIO_ERR_STATE, "two frames already in flight: call rsvio_tracker_sync");
    DeviceScope dev_scope;
    CU_TRY(t, dev_scope.enter(t->cfg.device));
    const size_t need = 2 * row * (size_t)h;
    if (t->raw_bytes < need) {
        CU_TRY(t, cudaStreamSynchronize(t->copy_st));
        cudaFree(t->d_raw);
        t->d_raw = nullptr;
        t->raw_bytes = 0;
        CU_TRY(t, cudaMalloc((void **)&t->d_raw, need));
        t->raw_bytes = need;
    }
    const int slot = t->next_slot;
    // raw upload and conversion on the copy stream: the conversion writes the slot the pyramid kernel reads
    CU_TRY(t, cudaStreamWaitEvent(t->copy_st, t->ev_in_free[slot], 0));
    if (t->profiling) cudaEventRecord(t->ev_up[slot][0], t->copy_st);
    const void *src[2] = {left, right};
    for (int cam = 0; cam < 2; ++cam) {
        uint8_t *raw = t->d_raw + (size_t)cam * row * h;
        CU_TRY(t, cudaMemcpy2DAsync(raw, row, src[cam], stride_bytes, row, (size_t)h, cudaMemcpyHostToDevice, t->copy_st));
        CU_TRY(t, convert_to_luma8(raw, row, w, h, pixel_format, t->d_in[slot] + (size_t)cam * t->img_bytes, t->copy_st));
    }
    t->stage_launches[0] += 4;
    t->host_launches += 2;
    if (t->profiling) cudaEventRecord(t->ev_up[slot][1], t->copy_st);
    t->up_valid[slot] = t->profiling;
    CU_TRY(t, cudaEventRecord(t->ev_h2d[slot], t->copy_st));
    t->slot_host[slot] = true;
    CU_TRY(t, cudaStreamWaitEvent(t->st, t->ev_h2d[slot], 0));
    if (t->pyr_st != t->st) CU_TRY(t, cudaStreamWaitEvent(t->pyr_st, t->ev_h2d[slot], 0));
    rsvio_status rs = enqueue_frame(t, slot, t->d_in[slot], t->d_in[slot] + t->img_bytes, (size_t)w, t->img_bytes);
    if (rs != RSVIO_OK) return rs;
    t->next_slot ^= 1;
    t->inflight += 1;
    // the caller may reuse its buffers when this returns: pageable sources have been staged by the runtime,
    // pinned ones are read by the DMA that the sync below waits for
    return rsvio_tracker_sync(t);
}

rsvio_status rsvio_tracker_process_mono(rsvio_tracker *t, const uint8_t *image, size_t stride) {
    if (!t) return RSVIO_ERR_INVALID_ARG;
    if (t->S != 1 || t->stereo) return fail(t, RSVIO_ERR_INVALID_ARG, "handle is not a 1-stream mono tracker");
    const uint8_t *l[1] = {image};
    return rsvio_tracker_process_batch(t, l, nullptr, stride);
}

// max_inflight = 1 for readers of the host-mapped result ring (the one frame still in flight
// writes the other slot), 0 for anything that touches the device tables.
static rsvio_status check_group(rsvio_tracker *t, int32_t stream, int32_t cam, int max_inflight = 0) {
    if (!t) return RSVIO_ERR_INVALID_ARG;
    if (stream < 0 || stream >= t->S || cam < 0 || cam > 1) return fail(t, RSVIO_ERR_INVALID_ARG, "bad stream/cam");
    if (t->inflight > max_inflight)
        return fail(t, RSVIO_ERR_STATE, "frames in flight: call rsvio_tracker_sync first");
    return RSVIO_OK;
}

rsvio_status rsvio_tracker_num_tracks(rsvio_tracker *t, int32_t stream, int32_t cam, int32_t *n) {
    rsvio_status rs = check_group(t, stream, cam, 1);
    if (rs != RSVIO_OK) return rs;
    if (!n) return RSVIO_ERR_INVALID_ARG;
    *n = t->result_slot < 0 ? 0 : t->h_cnt[t->result_slot][cam * t->S + stream];
    return RSVIO_OK;
}

rsvio_status rsvio_tracker_get_tracks(rsvio_tracker *t, int32_t stream, int32_t cam, uint64_t *ids, float *xy,
                                      size_t cap, size_t *n) {
    rsvio_status rs = check_group(t, stream, cam, 1);
    if (rs != RSVIO_OK) return rs;
    if (!n) return RSVIO_ERR_INVALID_ARG;
    const int g = cam * t->S + stream;
    const int cnt = t->result_slot < 0 ? 0 : t->h_cnt[t->result_slot][g];
    *n = (size_t)cnt;
    if ((size_t)cnt > cap) return fail(t, RSVIO_ERR_CAPACITY, "output arrays too small");
    if (cnt > 0 && (!ids || !xy)) return RSVIO_ERR_INVALID_ARG;
    if (cnt <= 0) return RSVIO_OK;  // also: no frame yet (result_slot < 0), nothing to index
    const OutRecord *r = t->h_out[t->result_slot] + (size_t)g * t->cap;
    for (int k = 0; k < cnt; ++k) {
        ids[k] = r[k].id;
        xy[2 * k] = r[k].x;
        xy[2 * k + 1] = r[k].y;
    }
    return RSVIO_OK;
}

rsvio_status rsvio_tracker_result_view(rsvio_tracker *t, const rsvio_track_record **records,
                                       const int32_t **counts, int32_t *groups, int32_t *capacity) {
    if (!t) return RSVIO_ERR_INVALID_ARG;
    if (t->inflight > 1) return fail(t, RSVIO_ERR_STATE, "two frames in flight: call rsvio_tracker_sync first");
    if (t->result_slot < 0) return fail(t, RSVIO_ERR_STATE, "no frame has been processed yet");
    static_assert(sizeof(rsvio_track_record) == sizeof(OutRecord), "record layout");
    if (records) *records = reinterpret_cast<const rsvio_track_record *>(t->h_out[t->result_slot]);
    if (counts) *counts = t->h_cnt[t->result_slot];
    if (groups) *groups = t->groups;
    if (capacity) *capacity = t->cap;
    return RSVIO_OK;
}

rsvio_status rsvio_tracker_set_camera(rsvio_tracker *t, int32_t cam, const rsvio_camera_model *model) {
    if (!t || !model || cam < 0 || cam > 1) return RSVIO_ERR_INVALID_ARG;
    if (model->model != RSVIO_CAMERA_NONE && model->model != RSVIO_CAMERA_PINHOLE_RADTAN &&
        model->model != RSVIO_CAMERA_EUCM)
        return fail(t, RSVIO_ERR_INVALID_ARG, "unknown camera model");
    if (model->model != RSVIO_CAMERA_NONE && !(model->params[0] != 0.0 && model->params[1] != 0.0))
        return fail(t, RSVIO_ERR_INVALID_ARG, "focal lengths must be non-zero");
    if (t->inflight > 0) return fail(t, RSVIO_ERR_STATE, "frames in flight: call rsvio_tracker_sync before set_camera");
    DeviceScope dev_scope;
    CU_TRY(t, dev_scope.enter(t->cfg.device));
    if (model->model != RSVIO_CAMERA_NONE && !t->h_uxy[0]) {
        const size_t bytes = (size_t)t->groups * t->cap * 2 * sizeof(float);
        for (int i = 0; i < kSlots; ++i) {
            CU_TRY(t, cudaHostAlloc((void **)&t->h_uxy[i], bytes, cudaHostAllocMapped));
            CU_TRY(t, cudaHostGetDevicePointer((void **)&t->d_uxy[i], t->h_uxy[i], 0));
            memset(t->h_uxy[i], 0, bytes);
        }
    }
    t->cam_model[cam].model = model->model;
    for (int k = 0; k < 9; ++k) t->cam_model[cam].p[k] = model->params[k];
    drop_graphs(t);  // the export kernel's parameters are baked into the captured frames
    return RSVIO_OK;
}

rsvio_status rsvio_tracker_undistorted_view(rsvio_tracker *t, const float **uxy, int32_t *groups, int32_t *capacity) {
    if (!t || !uxy) return RSVIO_ERR_INVALID_ARG;
    if (t->inflight > 1) return fail(t, RSVIO_ERR_STATE, "two frames in flight: call rsvio_tracker_sync first");
    if (t->result_slot < 0) return fail(t, RSVIO_ERR_STATE, "no frame has been processed yet");
    const bool on = t->cam_model[0].model != RSVIO_CAMERA_NONE || t->cam_model[1].model != RSVIO_CAMERA_NONE;
    *uxy = on ? t->h_uxy[t->result_slot] : nullptr;
    if (groups) *groups = t->groups;
    if (capacity) *capacity = t->cap;
    return RSVIO_OK;
}

rsvio_status rsvio_tracker_get_transforms(rsvio_tracker *t, int32_t stream, int32_t cam, uint64_t *ids, float *T6,
                                          size_t cap, size_t *n) {
    rsvio_status rs = check_group(t, stream, cam);
    if (rs != RSVIO_OK) return rs;
    if (!n) return RSVIO_ERR_INVALID_ARG;
    DeviceScope dev_scope;
    CU_TRY(t, dev_scope.enter(t->cfg.device));
    const int g = cam * t->S + stream;
    const TableView &tv = t->tab[t->tcur];
    int cnt = 0;
    CU_TRY(t, cudaMemcpyAsync(&cnt, tv.counts + g, sizeof(int), cudaMemcpyDeviceToHost, t->st));
    CU_TRY(t, cudaStreamSynchronize(t->st));
    *n = (size_t)cnt;
    if ((size_t)cnt > cap) return fail(t, RSVIO_ERR_CAPACITY, "output arrays too small");
    if (cnt > 0) {
        if (!ids || !T6) return RSVIO_ERR_INVALID_ARG;
        CU_TRY(t, cudaMemcpyAsync(ids, tv.ids + (size_t)g * t->cap, sizeof(uint64_t) * cnt, cudaMemcpyDeviceToHost, t->st));
        CU_TRY(t, cudaMemcpyAsync(T6, tv.T6 + (size_t)g * t->cap * 6, sizeof(float) * 6 * cnt, cudaMemcpyDeviceToHost, t->st));
        CU_TRY(t, cudaStreamSynchronize(t->st));
    }
    return RSVIO_OK;
}

rsvio_status rsvio_tracker_remove_ids(rsvio_tracker *t, int32_t stream, const uint64_t *ids, size_t n) {
    rsvio_status rs = check_group(t, stream, 0);
    if (rs != RSVIO_OK) return rs;
    if (n == 0) return RSVIO_OK;
    if (!ids) return RSVIO_ERR_INVALID_ARG;
    if (t->inflight > 0) return fail(t, RSVIO_ERR_STATE, "frames in flight: call rsvio_tracker_sync before remove_ids");
    DeviceScope dev_scope;
    CU_TRY(t, dev_scope.enter(t->cfg.device));
    if ((int)n > t->rm_cap) {
        cudaFree(t->d_rm);
        t->d_rm = nullptr;
        CU_TRY(t, dmalloc(&t->d_rm, n));
        t->rm_cap = (int)n;
    }
    CU_TRY(t, cudaMemcpyAsync(t->d_rm, ids, sizeof(uint64_t) * n, cudaMemcpyHostToDevice, t->st));
    CU_TRY(t, remove_ids(t->tab[t->tcur], t->tab[t->tcur ^ 1], t->groups, t->cap, t->S, stream, t->d_rm, (int)n, t->st));
    t->host_launches += 1;
    t->tcur ^= 1;
    t->surv_valid = false;  // rows moved: the next frame tracks the whole table in one launch
    // refresh the host-visible result of the affected groups
    const int slot = t->result_slot < 0 ? 0 : t->result_slot;
    const TableView &tv = t->tab[t->tcur];
    std::vector<int> cnt(t->groups);
    CU_TRY(t, cudaMemcpyAsync(cnt.data(), tv.counts, sizeof(int) * t->groups, cudaMemcpyDeviceToHost, t->st));
    CU_TRY(t, cudaStreamSynchronize(t->st));
    for (int cam = 0; cam < 2; ++cam) {
        const int g = cam * t->S + stream;
        std::vector<unsigned long long> hid(cnt[g] > 0 ? cnt[g] : 1);
        std::vector<float> hT((size_t)(cnt[g] > 0 ? cnt[g] : 1) * 6);
        if (cnt[g] > 0) {
            CU_TRY(t, cudaMemcpyAsync(hid.data(), tv.ids + (size_t)g * t->cap, sizeof(uint64_t) * cnt[g], cudaMemcpyDeviceToHost, t->st));
            CU_TRY(t, cudaMemcpyAsync(hT.data(), tv.T6 + (size_t)g * t->cap * 6, sizeof(float) * 6 * cnt[g], cudaMemcpyDeviceToHost, t->st));
            CU_TRY(t, cudaStreamSynchronize(t->st));
        }
        OutRecord *r = t->h_out[slot] + (size_t)g * t->cap;
        for (int k = 0; k < cnt[g]; ++k) {
            r[k].id = hid[k];
            r[k].x = hT[(size_t)k * 6 + 2];
            r[k].y = hT[(size_t)k * 6 + 5];
            if (t->h_uxy[slot] && t->cam_model[cam].model != RSVIO_CAMERA_NONE) {
                double ux, uy;
                unproject_pixel(t->cam_model[cam], (double)r[k].x, (double)r[k].y, ux, uy);
                float *u = t->h_uxy[slot] + ((size_t)g * t->cap + k) * 2;
                u[0] = (float)ux;
                u[1] = (float)uy;
            }
        }
        t->h_cnt[slot][g] = cnt[g];
    }
    t->result_slot = slot;
    return RSVIO_OK;
}

rsvio_status rsvio_tracker_counters(rsvio_tracker *t, uint64_t out[6]) {
    if (!t || !out) return RSVIO_ERR_INVALID_ARG;
    DeviceScope dev_scope;
    CU_TRY(t, dev_scope.enter(t->cfg.device));
    unsigned long long c[CNT_N];
    CU_TRY(t, cudaMemcpyAsync(c, t->d_counters, sizeof(c), cudaMemcpyDeviceToHost, t->st));
    CU_TRY(t, cudaStreamSynchronize(t->st));
    out[0] = c[CNT_TRACKS];
    out[1] = c[CNT_FRAMES];
    out[2] = c[CNT_ITERS];
    out[3] = c[CNT_TEMPLATES];
    out[4] = t->host_launches;
    out[5] = c[CNT_DETECTED];
    return RSVIO_OK;
}

rsvio_status rsvio_tracker_overflow_frames(rsvio_tracker *t, uint64_t *n) {
    if (!t || !n) return RSVIO_ERR_INVALID_ARG;
    *n = t->overflow_frames;
    return RSVIO_OK;
}

rsvio_status rsvio_tracker_set_profiling(rsvio_tracker *t, int32_t enabled) {
    if (!t) return RSVIO_ERR_INVALID_ARG;
    if (t->inflight > 0) return fail(t, RSVIO_ERR_STATE, "frames in flight");
    t->profiling = enabled != 0;
    return RSVIO_OK;
}

rsvio_status rsvio_tracker_stage_times(rsvio_tracker *t, float ms[RSVIO_N_STAGES], uint64_t launches[RSVIO_N_STAGES],
                                       int32_t reset) {
    if (!t) return RSVIO_ERR_INVALID_ARG;
    for (int k = 0; k < RSVIO_N_STAGES; ++k) {
        if (ms) ms[k] = (float)t->stage_ms[k];
        if (launches) launches[k] = t->stage_launches[k];
        if (reset) {
            t->stage_ms[k] = 0.0;
            t->stage_launches[k] = 0;
        }
    }
    return RSVIO_OK;
}

// ---- stateless entry points -------------------------------------------------------------

size_t rsvio_pyramid_bytes(int32_t width, int32_t height, int32_t levels) {
    size_t o = 0;
    for (int i = 0; i < levels; ++i) o += (size_t)(width >> i) * (size_t)(height >> i);
    return o;
}

#define CU_TRY0(expr)                                                                                   \
    do {                                                                                                \
        cudaError_t e__ = (expr);                                                                       \
        if (e__ != cudaSuccess) {                                                                       \
            fprintf(stderr, "rsvio: CUDA error %d (%s) at api.cu:%d: %s\n", (int)e__,                   \
                    cudaGetErrorString(e__), __LINE__, #expr);                                          \
            cudaGetLastError(); /* reported: clear the pending error */                                 \
            rs = RSVIO_ERR_CUDA;                                                                        \
            goto done;                                                                                  \
        }                                                                                               \
    } while (0)

// The stateless calls run on a private non-blocking stream (never the legacy default stream, which would
// serialise with every blocking stream of the host application and stall its tracker handles) and wait for
// that stream only.  Copies are stream-ordered on it as well.
struct CallStream {
    cudaStream_t s = nullptr;
    cudaError_t create() { return cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking); }
    ~CallStream() {
        if (s) cudaStreamDestroy(s);
    }
};
#define H2D(dst, src, bytes) cudaMemcpyAsync((dst), (src), (bytes), cudaMemcpyHostToDevice, cs.s)
#define D2H(dst, src, bytes) cudaMemcpyAsync((dst), (src), (bytes), cudaMemcpyDeviceToHost, cs.s)

static bool pyr_args_ok(int32_t n_images, int32_t width, int32_t height, size_t stride, size_t pitch, int32_t levels) {
    return n_images >= 1 && width >= 1 && height >= 1 && width <= 65535 && height <= 65535 && levels >= 1 &&
           levels <= RSVIO_MAX_LEVELS && (width >> (levels - 1)) >= 1 && (height >> (levels - 1)) >= 1 &&
           stride >= (size_t)width && (n_images == 1 || pitch >= stride * (size_t)(height - 1) + (size_t)width);
}

rsvio_status rsvio_build_pyramid_u8_device(int32_t device, const void *d_images, int32_t n_images, int32_t width,
                                           int32_t height, size_t stride, size_t pitch, int32_t levels, void *d_out,
                                           int32_t repeats, float *kernel_ms) {
    if (!d_images || !d_out || !pyr_args_ok(n_images, width, height, stride, pitch, levels) || repeats < 1)
        return RSVIO_ERR_INVALID_ARG;
    rsvio_status rs = RSVIO_OK;
    DeviceScope dev_scope;
    PyrPlan *plan = nullptr;
    cudaStream_t st = nullptr;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    CU_TRY0(dev_scope.enter(device));
    CU_TRY0(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    CU_TRY0(pyr_plan_create(width, height, levels, st, &plan));
    CU_TRY0(pyr_plan_reserve(plan, n_images));
    CU_TRY0(cudaEventCreate(&e0));
    CU_TRY0(cudaEventCreate(&e1));
    CU_TRY0(cudaEventRecord(e0, st));
    for (int r = 0; r < repeats; ++r)
        CU_TRY0(pyr_build(plan, (const uint8_t *)d_images, (const uint8_t *)d_images, n_images, n_images, stride, pitch,
                          (uint8_t *)d_out, st, nullptr));
    CU_TRY0(cudaEventRecord(e1, st));
    CU_TRY0(cudaStreamSynchronize(st));
    if (kernel_ms) {
        float ms = 0.0f;
        CU_TRY0(cudaEventElapsedTime(&ms, e0, e1));
        *kernel_ms = ms / (float)repeats;
    }
done:
    if (e0) cudaEventDestroy(e0);
    if (e1) cudaEventDestroy(e1);
    pyr_plan_destroy(plan);
    if (st) cudaStreamDestroy(st);
    return rs;
}

// internal (aligned) pyramid size, for callers that allocate d_out of the device variant
size_t rsvio_pyramid_device_bytes(int32_t width, int32_t height, int32_t levels) {
    return make_pyr_desc(width, height, levels).pyr_stride;
}

rsvio_status rsvio_build_pyramid_u8(int32_t device, const uint8_t *images, int32_t n_images, int32_t width,
                                    int32_t height, size_t stride, size_t pitch, int32_t levels, uint8_t *out) {
    if (!images || !out || !pyr_args_ok(n_images, width, height, stride, pitch, levels)) return RSVIO_ERR_INVALID_ARG;
    rsvio_status rs = RSVIO_OK;
    DeviceScope dev_scope;
    CallStream cs;
    const PyrDesc pd = make_pyr_desc(width, height, levels);
    const size_t dense = rsvio_pyramid_bytes(width, height, levels);
    uint8_t *d_in = nullptr, *d_out = nullptr;
    const size_t in_bytes = (size_t)(n_images - 1) * pitch + stride * (size_t)(height - 1) + (size_t)width;
    CU_TRY0(dev_scope.enter(device));
    CU_TRY0(cs.create());
    CU_TRY0(cudaMalloc((void **)&d_in, in_bytes));
    CU_TRY0(cudaMalloc((void **)&d_out, (size_t)n_images * pd.pyr_stride));
    CU_TRY0(H2D(d_in, images, in_bytes));
    CU_TRY0(cudaStreamSynchronize(cs.s));  // the device variant launches on a stream of its own
    rs = rsvio_build_pyramid_u8_device(device, d_in, n_images, width, height, stride, pitch, levels, d_out, 1, nullptr);
    if (rs != RSVIO_OK) goto done;
    for (int j = 0; j < n_images; ++j) {
        size_t o = 0;
        for (int i = 0; i < levels; ++i) {
            const size_t nb = (size_t)pd.lv[i].w * pd.lv[i].h;
            CU_TRY0(D2H(out + (size_t)j * dense + o, d_out + (size_t)j * pd.pyr_stride + pd.lv[i].off, nb));
            o += nb;
        }
    }
    CU_TRY0(cudaStreamSynchronize(cs.s));
done:
    cudaFree(d_in);
    cudaFree(d_out);
    return rs;
}

rsvio_status rsvio_track_points(int32_t device, const uint8_t *pyr_a, const uint8_t *pyr_b, int32_t width,
                                int32_t height, int32_t levels, int32_t n, const float *in_T6, int32_t max_iterations,
                                float convergence_threshold, float *out_T6, int32_t *status, int32_t *iters) {
    if (!pyr_a || !pyr_b || n < 0 || !pyr_args_ok(1, width, height, (size_t)width, 0, levels) || max_iterations < 0 ||
        !std::isfinite(convergence_threshold))
        return RSVIO_ERR_INVALID_ARG;
    if (n == 0) return RSVIO_OK;
    if (!in_T6 || !out_T6 || !status) return RSVIO_ERR_INVALID_ARG;
    rsvio_status rs = RSVIO_OK;
    DeviceScope dev_scope;
    CallStream cs;
    const PyrDesc pd = make_pyr_desc(width, height, levels);
    uint8_t *d_pyr = nullptr;
    float *d_in = nullptr, *d_out = nullptr;
    int32_t *d_status = nullptr, *d_iters = nullptr;
    int *d_count = nullptr;
    CU_TRY0(dev_scope.enter(device));
    CU_TRY0(cs.create());
    CU_TRY0(cudaMalloc((void **)&d_pyr, 2 * (size_t)pd.pyr_stride));
    {
        size_t o = 0;
        for (int i = 0; i < levels; ++i) {
            const size_t nb = (size_t)pd.lv[i].w * pd.lv[i].h;
            CU_TRY0(H2D(d_pyr + pd.lv[i].off, pyr_a + o, nb));
            CU_TRY0(H2D(d_pyr + pd.pyr_stride + pd.lv[i].off, pyr_b + o, nb));
            o += nb;
        }
    }
    CU_TRY0(cudaMalloc((void **)&d_in, sizeof(float) * 6 * n));
    CU_TRY0(cudaMalloc((void **)&d_out, sizeof(float) * 6 * n));
    CU_TRY0(cudaMalloc((void **)&d_status, sizeof(int32_t) * n));
    CU_TRY0(cudaMalloc((void **)&d_iters, sizeof(int32_t) * 2 * n));
    CU_TRY0(cudaMalloc((void **)&d_count, sizeof(int)));
    CU_TRY0(H2D(d_in, in_T6, sizeof(float) * 6 * n));
    CU_TRY0(H2D(d_count, &n, sizeof(int)));
    {
        LkBatch b{};
        b.pyr_a = d_pyr;
        b.pyr_b = d_pyr + pd.pyr_stride;
        b.counts = d_count;
        b.in_T6 = d_in;
        b.out_T6 = d_out;
        b.status = d_status;
        b.iters = d_iters;
        b.counters = nullptr;
        b.work_counter = nullptr;  // one group: static striding is already balanced
        b.groups = 1;
        b.cap = n;
        b.max_it = max_iterations;
        b.thr = convergence_threshold;
        CU_TRY0(lk_track(pd, b, cs.s));
    }
    CU_TRY0(cudaStreamSynchronize(cs.s));
    CU_TRY0(D2H(out_T6, d_out, sizeof(float) * 6 * n));
    CU_TRY0(D2H(status, d_status, sizeof(int32_t) * n));
    if (iters) CU_TRY0(D2H(iters, d_iters, sizeof(int32_t) * 2 * n));
    CU_TRY0(cudaStreamSynchronize(cs.s));
done:
    cudaFree(d_pyr);
    cudaFree(d_in);
    cudaFree(d_out);
    cudaFree(d_status);
    cudaFree(d_iters);
    cudaFree(d_count);
    return rs;
}

rsvio_status rsvio_detect_keypoints(int32_t device, const uint8_t *image, int32_t width, int32_t height, size_t stride,
                                    uint32_t grid_size, const uint32_t *cur_xy, int32_t n_cur, uint32_t *out_xy,
                                    float *out_score, size_t cap, size_t *n) {
    if (!image || !n || width < 8 || height < 8 || width > 65535 || height > 65535 || stride < (size_t)width ||
        grid_size < 7 || grid_size > 1024 || n_cur < 0 || (n_cur > 0 && !cur_xy))
        return RSVIO_ERR_INVALID_ARG;
    rsvio_status rs = RSVIO_OK;
    DeviceScope dev_scope;
    CallStream cs;
    const GridDesc gd = make_grid_desc(width, height, grid_size);
    const int cells = gd.ncx * gd.ncy;
    *n = 0;
    if (cells <= 0) return RSVIO_OK;
    uint8_t *d_img = nullptr, *d_occ = nullptr;
    uint32_t *d_xy = nullptr, *d_cell = nullptr;
    float *d_score = nullptr, *d_T6 = nullptr, *d_nscore = nullptr;
    int *d_cnt = nullptr, *d_list_n = nullptr;  // d_list_n: [0,1] list counters, [2] listed cells, [3] finished cells
    uint32_t *d_list = nullptr;
    int cnt = 0;
    std::vector<float> hT, hs;
    CU_TRY0(dev_scope.enter(device));
    CU_TRY0(cs.create());
    CU_TRY0(cudaMalloc((void **)&d_img, (size_t)width * height));
    CU_TRY0(cudaMemcpy2DAsync(d_img, (size_t)width, image, stride, (size_t)width, (size_t)height, cudaMemcpyHostToDevice, cs.s));
    CU_TRY0(cudaMalloc((void **)&d_occ, (size_t)gd.occ_rows * gd.occ_cols));
    CU_TRY0(cudaMemsetAsync(d_occ, 0, (size_t)gd.occ_rows * gd.occ_cols, cs.s));
    if (n_cur > 0) {
        CU_TRY0(cudaMalloc((void **)&d_xy, sizeof(uint32_t) * 2 * n_cur));
        CU_TRY0(H2D(d_xy, cur_xy, sizeof(uint32_t) * 2 * n_cur));
        CU_TRY0(mark_occupancy_xy(gd, d_xy, n_cur, d_occ, cs.s));
    }
    CU_TRY0(cudaMalloc((void **)&d_cell, sizeof(uint32_t) * cells));
    CU_TRY0(cudaMalloc((void **)&d_score, sizeof(float) * cells));
    CU_TRY0(cudaMalloc((void **)&d_T6, sizeof(float) * 6 * cells));
    CU_TRY0(cudaMalloc((void **)&d_nscore, sizeof(float) * cells));
    CU_TRY0(cudaMalloc((void **)&d_cnt, sizeof(int)));
    CU_TRY0(cudaMalloc((void **)&d_list, sizeof(uint32_t) * cells));
    CU_TRY0(cudaMalloc((void **)&d_list_n, 4 * sizeof(int)));
    CU_TRY0(cudaMemsetAsync(d_list_n, 0, 4 * sizeof(int), cs.s));
    CU_TRY0(cudaMemsetAsync(d_cnt, 0, sizeof(int), cs.s));  // stays 0 when no cell is empty
    {
        CellWork cw{};
        cw.cell_pt = d_cell;
        cw.cell_score = d_score;
        cw.list = d_list;
        cw.list_count = d_list_n;
        cw.next_count = d_list_n + 1;
        cw.n_empty = d_list_n + 2;
        cw.done = d_list_n + 3;
        cw.new_count = d_cnt;
        CU_TRY0(detect_cells(gd, d_img, (size_t)width, 0, 1, d_occ, cw, d_T6, d_nscore, cells, nullptr, true, false, cs.s));
    }
    CU_TRY0(D2H(&cnt, d_cnt, sizeof(int)));
    CU_TRY0(cudaStreamSynchronize(cs.s));
    *n = (size_t)cnt;
    if ((size_t)cnt > cap) {
        rs = RSVIO_ERR_CAPACITY;
        goto done;
    }
    if (cnt > 0) {
        if (!out_xy) {
            rs = RSVIO_ERR_INVALID_ARG;
            goto done;
        }
        hT.resize((size_t)cnt * 6);
        hs.resize((size_t)cnt);
        CU_TRY0(D2H(hT.data(), d_T6, sizeof(float) * 6 * cnt));
        CU_TRY0(D2H(hs.data(), d_nscore, sizeof(float) * cnt));
        CU_TRY0(cudaStreamSynchronize(cs.s));
        for (int k = 0; k < cnt; ++k) {
            out_xy[2 * k] = (uint32_t)hT[(size_t)k * 6 + 2];
            out_xy[2 * k + 1] = (uint32_t)hT[(size_t)k * 6 + 5];
            if (out_score) out_score[k] = hs[k];
        }
    }
done:
    cudaFree(d_img);
    cudaFree(d_occ);
    cudaFree(d_xy);
    cudaFree(d_cell);
    cudaFree(d_score);
    cudaFree(d_T6);
    cudaFree(d_nscore);
    cudaFree(d_cnt);
    cudaFree(d_list);
    cudaFree(d_list_n);
    return rs;
}

rsvio_status rsvio_se2_exp(int32_t device, const float *twists, int32_t n, float *out9) {
    if (n < 0 || (n > 0 && (!twists || !out9))) return RSVIO_ERR_INVALID_ARG;
    if (n == 0) return RSVIO_OK;
    rsvio_status rs = RSVIO_OK;
    DeviceScope dev_scope;
    CallStream cs;
    float *d_in = nullptr, *d_out = nullptr;
    CU_TRY0(dev_scope.enter(device));
    CU_TRY0(cs.create());
    CU_TRY0(cudaMalloc((void **)&d_in, sizeof(float) * 3 * n));
    CU_TRY0(cudaMalloc((void **)&d_out, sizeof(float) * 9 * n));
    CU_TRY0(H2D(d_in, twists, sizeof(float) * 3 * n));
    CU_TRY0(se2_exp_batch(d_in, n, d_out, cs.s));
    CU_TRY0(cudaStreamSynchronize(cs.s));
    CU_TRY0(D2H(out9, d_out, sizeof(float) * 9 * n));
    CU_TRY0(cudaStreamSynchronize(cs.s));
done:
    cudaFree(d_in);
    cudaFree(d_out);
    return rs;
}

rsvio_status rsvio_microbench_fp32(int32_t device, float *ffma_tflops, float *fmul_fadd_tflops) {
    if (!ffma_tflops || !fmul_fadd_tflops) return RSVIO_ERR_INVALID_ARG;
    rsvio_status rs = RSVIO_OK;
    DeviceScope dev_scope;
    CallStream cs;
    CU_TRY0(dev_scope.enter(device));
    CU_TRY0(cs.create());
    CU_TRY0(fp32_rate(ffma_tflops, fmul_fadd_tflops, cs.s));
done:
    return rs;
}

rsvio_status rsvio_unproject_points(int32_t device, const rsvio_camera_model *model, const float *xy, int32_t n,
                                    float *out_xy) {
    if (!model || !xy || !out_xy || n < 0) return RSVIO_ERR_INVALID_ARG;
    if (model->model != RSVIO_CAMERA_PINHOLE_RADTAN && model->model != RSVIO_CAMERA_EUCM) return RSVIO_ERR_INVALID_ARG;
    if (n == 0) return RSVIO_OK;
    rsvio_status rs = RSVIO_OK;
    DeviceScope dev_scope;
    CallStream cs;
    float *d_in = nullptr, *d_out = nullptr;
    CameraModelDev cm{};
    cm.model = model->model;
    for (int k = 0; k < 9; ++k) cm.p[k] = model->params[k];
    CU_TRY0(dev_scope.enter(device));
    CU_TRY0(cs.create());
    CU_TRY0(cudaMalloc((void **)&d_in, sizeof(float) * 2 * (size_t)n));
    CU_TRY0(cudaMalloc((void **)&d_out, sizeof(float) * 2 * (size_t)n));
    CU_TRY0(H2D(d_in, xy, sizeof(float) * 2 * (size_t)n));
    CU_TRY0(unproject_points(cm, d_in, n, d_out, cs.s));
    CU_TRY0(D2H(out_xy, d_out, sizeof(float) * 2 * (size_t)n));
    CU_TRY0(cudaStreamSynchronize(cs.s));
done:
    cudaFree(d_in);
    cudaFree(d_out);
    return rs;
}

rsvio_status rsvio_convert_to_luma8(int32_t device, const void *src, int32_t width, int32_t height, size_t stride_bytes,
                                    int32_t pixel_format, uint8_t *out) {
    const int bpp = pixel_bytes(pixel_format);
    if (!src || !out || width < 1 || height < 1 || bpp == 0 || stride_bytes < (size_t)width * bpp)
        return RSVIO_ERR_INVALID_ARG;
    rsvio_status rs = RSVIO_OK;
    DeviceScope dev_scope;
    CallStream cs;
    uint8_t *d_src = nullptr, *d_dst = nullptr;
    const size_t row = (size_t)width * bpp;
    CU_TRY0(dev_scope.enter(device));
    CU_TRY0(cs.create());
    CU_TRY0(cudaMalloc((void **)&d_src, row * (size_t)height));
    CU_TRY0(cudaMalloc((void **)&d_dst, (size_t)width * height));
    CU_TRY0(cudaMemcpy2DAsync(d_src, row, src, stride_bytes, row, (size_t)height, cudaMemcpyHostToDevice, cs.s));
    CU_TRY0(convert_to_luma8(d_src, row, width, height, pixel_format, d_dst, cs.s));
    CU_TRY0(D2H(out, d_dst, (size_t)width * height));
    CU_TRY0(cudaStreamSynchronize(cs.s));
done:
    cudaFree(d_src);
    cudaFree(d_dst);
    return rs;
}

rsvio_status rsvio_b_pyramid_dims(int32_t width, int32_t height, int32_t levels, double ratio, int32_t *widths,
                                  int32_t *heights) {
    if (width < 1 || height < 1 || levels < 1 || levels > RSVIO_MAX_LEVELS || !(ratio > 1.0) || !widths || !heights)
        return RSVIO_ERR_INVALID_ARG;
    double f = 1.0;
    for (int l = 0; l < levels; ++l) {
        // (w0 as f64 / ratio.powi(l)).round() as u32 -- round half away from zero
        widths[l] = (int32_t)std::floor((double)width / f + 0.5);
        heights[l] = (int32_t)std::floor((double)height / f + 0.5);
        if (widths[l] < 1 || heights[l] < 1) return RSVIO_ERR_INVALID_ARG;
        f *= ratio;
    }
    return RSVIO_OK;
}

rsvio_status rsvio_b_build_pyramid(int32_t device, const float *image, int32_t width, int32_t height, int32_t levels,
                                   double ratio, float *const *out_levels) {
    int32_t w[RSVIO_MAX_LEVELS], h[RSVIO_MAX_LEVELS];
    if (!image || !out_levels) return RSVIO_ERR_INVALID_ARG;
    rsvio_status rs = rsvio_b_pyramid_dims(width, height, levels, ratio, w, h);
    if (rs != RSVIO_OK) return rs;
    for (int l = 0; l < levels; ++l)
        if (!out_levels[l]) return RSVIO_ERR_INVALID_ARG;
    DeviceScope dev_scope;
    CallStream cs;
    float *d_lv[RSVIO_MAX_LEVELS] = {};
    float *d_tmp = nullptr;
    CU_TRY0(dev_scope.enter(device));
    CU_TRY0(cs.create());
    for (int l = 0; l < levels; ++l) CU_TRY0(cudaMalloc((void **)&d_lv[l], sizeof(float) * (size_t)w[l] * h[l]));
    CU_TRY0(cudaMalloc((void **)&d_tmp, sizeof(float) * (size_t)w[0] * h[0]));
    CU_TRY0(H2D(d_lv[0], image, sizeof(float) * (size_t)w[0] * h[0]));
    for (int l = 1; l < levels; ++l)
        CU_TRY0(resize_triangle_f32(d_lv[l - 1], w[l - 1], h[l - 1], d_tmp, d_lv[l], w[l], h[l], cs.s));
    CU_TRY0(cudaStreamSynchronize(cs.s));
    for (int l = 0; l < levels; ++l)
        CU_TRY0(D2H(out_levels[l], d_lv[l], sizeof(float) * (size_t)w[l] * h[l]));
    CU_TRY0(cudaStreamSynchronize(cs.s));
done:
    for (int l = 0; l < RSVIO_MAX_LEVELS; ++l) cudaFree(d_lv[l]);
    cudaFree(d_tmp);
    return rs;
}

rsvio_status rsvio_b_track_points(int32_t device, const float *const *pyr0, const float *const *pyr1,
                                  const int32_t *widths, const int32_t *heights, int32_t levels,
                                  const float *points_xy, int32_t n, int32_t max_iteration, float lm_lambda,
                                  float *out_T4, int32_t *status) {
    return rsvio_b_track_points_cost(device, pyr0, pyr1, widths, heights, levels, points_xy, n, max_iteration, lm_lambda,
                                     RSVIO_B_COST_SSD, out_T4, status);
}

rsvio_status rsvio_b_track_points_cost(int32_t device, const float *const *pyr0, const float *const *pyr1,
                                       const int32_t *widths, const int32_t *heights, int32_t levels,
                                       const float *points_xy, int32_t n, int32_t max_iteration, float lm_lambda,
                                       int32_t matching_cost, float *out_T4, int32_t *status) {
    if (!pyr0 || !pyr1 || !widths || !heights || levels < 1 || levels > RSVIO_MAX_LEVELS || !points_xy || n < 0 ||
        max_iteration < 0 || !out_T4 || !status || (matching_cost != RSVIO_B_COST_SSD && matching_cost != RSVIO_B_COST_LSSD))
        return RSVIO_ERR_INVALID_ARG;
    for (int l = 0; l < levels; ++l)
        if (!pyr0[l] || !pyr1[l] || widths[l] < 1 || heights[l] < 1) return RSVIO_ERR_INVALID_ARG;
    if (n == 0) return RSVIO_OK;
    rsvio_status rs = RSVIO_OK;
    DeviceScope dev_scope;
    CallStream cs;
    float *d_lv[2][RSVIO_MAX_LEVELS] = {};
    float *d_pts = nullptr, *d_T = nullptr;
    int32_t *d_st = nullptr;
    int w[RSVIO_MAX_LEVELS], h[RSVIO_MAX_LEVELS];
    CU_TRY0(dev_scope.enter(device));
    CU_TRY0(cs.create());
    for (int l = 0; l < levels; ++l) {
        w[l] = widths[l];
        h[l] = heights[l];
        const size_t bytes = sizeof(float) * (size_t)w[l] * h[l];
        for (int k = 0; k < 2; ++k) {
            CU_TRY0(cudaMalloc((void **)&d_lv[k][l], bytes));
            CU_TRY0(H2D(d_lv[k][l], (k == 0 ? pyr0 : pyr1)[l], bytes));
        }
    }
    CU_TRY0(cudaMalloc((void **)&d_pts, sizeof(float) * 2 * (size_t)n));
    CU_TRY0(cudaMalloc((void **)&d_T, sizeof(float) * 4 * (size_t)n));
    CU_TRY0(cudaMalloc((void **)&d_st, sizeof(int32_t) * (size_t)n));
    CU_TRY0(H2D(d_pts, points_xy, sizeof(float) * 2 * (size_t)n));
    CU_TRY0(trackb_points(d_lv[0], d_lv[1], w, h, levels, d_pts, n, max_iteration, lm_lambda, matching_cost, d_T, d_st, cs.s));
    CU_TRY0(cudaStreamSynchronize(cs.s));
    CU_TRY0(D2H(out_T4, d_T, sizeof(float) * 4 * (size_t)n));
    CU_TRY0(D2H(status, d_st, sizeof(int32_t) * (size_t)n));
    CU_TRY0(cudaStreamSynchronize(cs.s));
done:
    for (int l = 0; l < RSVIO_MAX_LEVELS; ++l)
        for (int k = 0; k < 2; ++k) cudaFree(d_lv[k][l]);
    cudaFree(d_pts);
    cudaFree(d_T);
    cudaFree(d_st);
    return rs;
}

}  // extern "C"
