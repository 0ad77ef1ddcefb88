#!/usr/bin/env python
"""bench.py -- stereo feature-tracker throughput on B200 (BASELINE.json metric).

Workload (config.workload): BASELINE.json configs[4] "C5": independent EuRoC-shape 752x480 stereo
streams, 3-level pyramid, grid 40 (216 cells), LK 20 iterations / 0.01, synthetic analytic texture
with known sub-pixel motion.  Every GPU (one process per GPU, no collective on the data path) runs
`--streams-per-gpu` streams (default 64) batched on one tracker handle; a *step* is one stereo frame
of every stream on the GPU: upload, 2 pyramids, temporal LK (both cameras), grid FAST replenishment,
stereo LK, id merge, packed result.  Frames come from a ring of 8 consecutive synthetic frames played
forwards then backwards, so the motion magnitude and the track population stay in steady state.

value   = feature tracks/s (one track = one forward+backward LK chain over all levels, accepted or
          not) with the frames already resident in HBM, whole job, max time over ranks.
e2e     = the same through the reference-facing API with HOST (pinned) images: the H2D copy of every
          step's frames and the read of the host-mapped result are inside the timed region.
roofline= the HBM-bound kernel (pyramid): algorithmic bytes (read w*h + write all levels per image)
          / average launch time from CUDA events inside this run, against MEASURED_PEAKS.json.
--impl reference: the reference's CPU algorithm (oracle port; the Rust crate cannot be built in this
          image) on all host threads, same metric/config, rank 0 only.
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (ROOT, os.path.join(ROOT, "rs-vio_b200")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import numpy as np  # noqa: E402

from rsvio_b200 import sharding  # noqa: E402
from rsvio_b200.synth import CONFIGS, StereoSequence  # noqa: E402

RING = 8
ALG_BYTES_C1_IMAGE = 752 * 480 + (752 * 480 + 376 * 240 + 188 * 120)  # read base + write 3 levels


def _render_stream(args):
    config, stream_id, ring = args
    seq = StereoSequence(config, stream_id=stream_id)
    out = np.empty((ring, 2, seq.params["h"], seq.params["w"]), np.uint8)
    for t in range(ring):
        out[t, 0], out[t, 1] = seq.frame(t)
    return stream_id, out


def generate_frames(config, stream_ids, ring, workers):
    """[ring][cam][stream][h][w] uint8, seeded per stream (rsvio_b200.synth)."""
    p = CONFIGS[config]
    # RSVIO_FRAME_CACHE=<dir> (profiling scripts that start this program several times on one box): render once
    cache = os.environ.get("RSVIO_FRAME_CACHE")
    cpath = None
    if cache and stream_ids:
        cpath = os.path.join(cache, f"frames_{config}_{stream_ids[0]}_{len(stream_ids)}_{ring}.npy")
        if os.path.exists(cpath):
            return np.load(cpath)
    out = np.empty((ring, 2, len(stream_ids), p["h"], p["w"]), np.uint8)
    jobs = [(config, sid, ring) for sid in stream_ids]
    if workers > 1:
        import multiprocessing as mp
        with mp.get_context("fork").Pool(workers) as pool:
            res = pool.map(_render_stream, jobs, chunksize=1)
    else:
        res = [_render_stream(j) for j in jobs]
    pos = {sid: k for k, sid in enumerate(stream_ids)}
    for sid, arr in res:
        out[:, :, pos[sid]] = arr
    if cpath:
        os.makedirs(cache, exist_ok=True)
        np.save(cpath, out)
    return out


def ring_index(i, ring=RING):
    """0,1,..,ring-1,ring-2,..,1,0,1,.. (forwards then backwards: consecutive frames stay 1 step apart)"""
    period = 2 * (ring - 1)
    k = i % period
    return k if k < ring else period - k


class ClockSampler(threading.Thread):
    """samples SM clock and throttle reasons of one GPU with NVML while the timed regions run"""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.active, self.stop_flag = index, [], set(), False, False
        self.max_mhz = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {
            "hw_slowdown": getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8),
            "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
            "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
            "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4),
            "hw_power_brake_slowdown": getattr(nv, "nvmlClocksEventReasonHwPowerBrakeSlowdown", 0x80),
        }
        while not self.stop_flag:
            if self.active:
                try:
                    self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                    try:
                        mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                    except Exception:
                        mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                    for k, bit in names.items():
                        if mask & bit:
                            self.reasons.add(k)
                except Exception:
                    pass
            time.sleep(0.005)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# ------------------------------------------------------------------------------------------------
# CPU legs (oracle port): cpu_baseline of the default run, and the whole --impl reference arm
# ------------------------------------------------------------------------------------------------
def cpu_run(config, frames, n_streams, n_frames, warm_frames, threads):
    """oracle StereoTracker per stream, streams spread over `threads` host threads (ctypes releases the
    GIL), frames [0, warm) untimed.  A single stream gets all threads inside the tracker instead (levels and
    points in parallel, as the reference's rayon calls).  Returns (tracks, stereo_frames, seconds)."""
    from concurrent.futures import ThreadPoolExecutor
    from oracle import oracle as orc
    orc.build()
    p = CONFIGS[config]
    inner = threads if n_streams == 1 else 1
    trackers = [orc.StereoTracker(p["levels"], p["grid"], p["max_it"], p["thr"], nthreads=inner) for _ in range(n_streams)]

    def advance(s, t0, t1):
        for t in range(t0, t1):
            k = ring_index(t, frames.shape[0])
            trackers[s].process_frame(frames[k, 0, s], frames[k, 1, s])

    with ThreadPoolExecutor(max_workers=max(1, min(threads, n_streams))) as ex:
        list(ex.map(lambda s: advance(s, 0, warm_frames), range(n_streams)))
        before = sum(t.counters()[0] for t in trackers)
        t0 = time.perf_counter()
        list(ex.map(lambda s: advance(s, warm_frames, warm_frames + n_frames), range(n_streams)))
        dt = time.perf_counter() - t0
    tracks = sum(t.counters()[0] for t in trackers) - before
    return tracks, n_streams * n_frames, dt


def run_reference(args, rank):
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    S = args.streams_per_gpu
    frames = generate_frames(args.config, list(range(S)), RING, max(1, min(16, cores)))
    # one step = one stereo frame of every stream, as in the GPU arm
    tracks, nframes, dt = cpu_run(args.config, frames, S, args.steps, args.warmup, cores)
    value = tracks / dt
    line = {
        "impl": "reference", "metric": "feature_tracks_per_sec", "value": value, "unit": "tracks/s",
        "stereo_frames_per_sec": nframes / dt, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, S, args.gpus),
        "cpu_baseline": {"value": value, "unit": "tracks/s", "cores": cores, "kind": "port",
                         "sample": f"{S} streams x {args.steps} frames after {args.warmup} warm-up frames, "
                                   f"oracle port (C++ -O3, f32, no FMA) one stream per thread on {cores} threads"},
        "e2e": {"value": value, "unit": "tracks/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_config(args, S, n_gpus):
    p = CONFIGS[args.config]
    total = args.total_streams if getattr(args, "total_streams", None) else S * n_gpus
    return {"workload": f"{args.config}: {S} independent EuRoC-shape {p['w']}x{p['h']} stereo streams per GPU, "
                        f"{p['levels']}-level pyramid, grid {p['grid']}, LK {p['max_it']} it / {p['thr']}",
            "streams_per_gpu": S, "total_streams": total, "frame_ring": RING,
            "l2": "per step the kernels touch previous+current pyramids and the input ring slot "
                  "(> 230 MB at 64 streams), larger than the 126 MB L2; no explicit flush",
            "parallelism": f"{n_gpus} x independent replicas (streams sharded, no data-path collective; "
                           "torch.distributed/NCCL carries only the timing barrier and the final counter reduction)",
            "pipeline": "one handle per GPU, <= 2 frames in flight; with frames resident in HBM the temporal LK of a frame is "
                        "launched in three parts (survivors of the previous compaction beside the previous frame's "
                        "detection / stereo LK / merge; newly merged points; helper), host-fed frames are upload-bound "
                        "and run unsplit; stage_ms_per_step / roofline are measured in profiling mode (one stream, "
                        "one temporal launch, CUDA events between the stages)"}


# ------------------------------------------------------------------------------------------------
def ncu_dram_traffic(same_workload):
    """dram__bytes_read.sum + dram__bytes_write.sum of the pyramid kernel, per launch, from the committed
    `ncu --set full` capture of this same workload (profiles/r2_pyr_fast_kernel.txt, else round 1's); None for
    other workloads.  A number taken under a profiler is evidence, never a bench value: only `traffic` uses it."""
    if not same_workload:
        return None, None
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    for name in ("r2_pyr_fast_kernel.txt", "r1_pyr_fast_kernel.txt"):
        path = os.path.join(ROOT, "profiles", name)
        if not os.path.exists(path):
            continue
        tot = 0.0
        for line in open(path):
            f = line.split()
            if len(f) >= 3 and f[0] in ("dram__bytes_read.sum", "dram__bytes_write.sum") and f[2] in scale:
                tot += float(f[1]) * scale[f[2]]
        if tot:
            return tot, f"profiles/{name} (ncu --set full, cold L2, one launch of 128 images)"
    return None, None


def bind_to_gpu_numa_node(local_rank):
    """Multi-GPU runs: keep this rank's threads -- and with them the first-touch pages of the pinned frame ring --
    on the CPUs NVML reports as local to its GPU, so that the uploads do not cross the socket interconnect.
    Returns the number of CPUs bound to, or None when NVML has no answer."""
    try:
        import pynvml
        pynvml.nvmlInit()
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        idx = int(vis.split(",")[local_rank]) if vis and vis.split(",")[local_rank].isdigit() else local_rank
        hdl = pynvml.nvmlDeviceGetHandleByIndex(idx)
        words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(hdl, words)
        cpus = {64 * i + b for i, wd in enumerate(mask) for b in range(64) if (wd >> b) & 1}
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return len(cpus)
    except Exception:
        pass
    return None


def sum_live_records(rec, cnt):
    """sum of x and y over the live records of every group (records[group][k], k < cnt[group]) in ONE masked
    reduction: a Python loop over the 128 groups of the bench handle took 0.86 ms per step -- the whole upload time of a
    step -- and left the submit of the next-but-one upload late (two frames in flight), i.e. gaps in the DMA stream"""
    cmax = int(cnt.max()) if cnt.size else 0
    if cmax <= 0:
        return 0.0
    live = np.arange(cmax, dtype=np.int32)[None, :] < cnt[:, None]
    return float(rec["x"][:, :cmax][live].sum(dtype=np.float64)) + float(rec["y"][:, :cmax][live].sum(dtype=np.float64))


class Runner:
    """One tracker handle on one GPU fed from a ring of synthetic frames: device-resident steps (`value`),
    host-frame steps through the reference-facing submit/sync calls (`e2e`), per-stage profiling."""

    def __init__(self, torch, rsvio_b200, config, frames_np, local_rank, stream, host_kind="wc"):
        import ctypes as C
        self.torch, self.rs, self.C = torch, rsvio_b200, C
        self.p = CONFIGS[config]
        self.w, self.h = self.p["w"], self.p["h"]
        self.S = frames_np.shape[2]
        self.ring = frames_np.shape[0]
        self.stream = stream
        dev = torch.device("cuda", local_rank)
        # Pinned frame ring [ring][cam][S][h][w].  "wc": rsvio_host_alloc_flags(WRITE_COMBINED) -- the CPU only writes
        # these pages, the DMA engine reads them without snooping CPU caches; "pinned": torch.pin_memory().
        self.host_kind = host_kind
        if host_kind == "wc":
            self.hostbuf = rsvio_b200.HostBuffer(frames_np.shape, write_combined=True)
            np.copyto(self.hostbuf.array, frames_np)
            base = self.hostbuf.ptr
        else:
            self.hostbuf = torch.from_numpy(frames_np).pin_memory()
            base = self.hostbuf.data_ptr()
        self.dev = torch.from_numpy(frames_np).to(dev)
        img = self.w * self.h
        self.img_bytes = img
        self.host_ptrs = []
        for k in range(self.ring):
            lp = (C.c_void_p * self.S)(*[base + ((k * 2 + 0) * self.S + s) * img for s in range(self.S)])
            rp = (C.c_void_p * self.S)(*[base + ((k * 2 + 1) * self.S + s) * img for s in range(self.S)])
            self.host_ptrs.append((lp, rp))
        with torch.cuda.stream(stream):
            self.tr = rsvio_b200.StereoPatchTracker(self.p["grid"], self.p["max_it"], self.p["thr"], width=self.w,
                                                    height=self.h, levels=self.p["levels"], n_streams=self.S,
                                                    device=local_rank, cuda_stream=stream.cuda_stream)
        self.read_sum = 0.0

    def close(self):
        self.tr.close()
        if self.host_kind == "wc":
            self.hostbuf.close()

    def _submit(self, kind, i):
        k = ring_index(i, self.ring)
        if kind == "device":
            self.tr.submit_device(self.dev[k, 0].data_ptr(), self.dev[k, 1].data_ptr(), self.w, self.img_bytes)
        else:
            lp, rp = self.host_ptrs[k]
            self.tr.submit_batch_ptr(lp, rp, self.w)

    def _consume(self):
        """read the WHOLE result of the step on the host (zero-copy view of host-mapped pinned memory)"""
        rec, cnt = self.tr.result_view()
        self.read_sum += sum_live_records(rec, cnt)
        return int(cnt.sum())

    def run(self, kind, first, n, consume=False):
        """n pipelined steps (<= 2 in flight); returns the number of records read"""
        recs = 0
        self._submit(kind, first)
        for i in range(first + 1, first + n):
            self._submit(kind, i)
            self.tr.sync()
            if consume:
                recs += self._consume()
        self.tr.sync()
        if consume:
            recs += self._consume()
        return recs

    def timed(self, kind, first, n, barrier, sampler=None):
        """(counter deltas, device ms between events on the handle's stream, wall s, records read)"""
        torch = self.torch
        barrier()
        c0 = self.tr.counters()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        if sampler is not None:
            sampler.active = True
        t0 = time.perf_counter()
        e0.record(self.stream)
        recs = self.run(kind, first, n, consume=(kind == "host"))
        e1.record(self.stream)
        e1.synchronize()
        wall = time.perf_counter() - t0
        if sampler is not None:
            sampler.active = False
        barrier()
        c1 = self.tr.counters()
        return {k: c1[k] - c0[k] for k in c1}, e0.elapsed_time(e1), wall, recs

    def stages(self, kind, first, n):
        """per-stage device ms per step: profiling mode puts everything on one stream with CUDA events between
        the stages, so a stage time is that stage's kernels alone on the device"""
        self.tr.set_profiling(True)
        self.run(kind, first, 3)           # warm-up in profiling mode (one stream, events between the stages)
        self.tr.stage_times(reset=True)
        first += 3
        self.run(kind, first, n)
        st = self.tr.stage_times(reset=True)
        self.tr.set_profiling(False)
        return {k: v[0] / n for k, v in st.items()}, {k: v[1] // n for k, v in st.items()}


def per_config_results(torch, rsvio_b200, local_rank, stream, cores, which, steps=24, warmup=6):
    """value / e2e / cpu_baseline of the single-stream BASELINE configs (C1..C4 and the estimator's shipped L = 6
    setting): one stereo stream through one handle, short runs, same definitions as the headline numbers."""
    out = []
    for config in which:
        try:
            p = CONFIGS[config]
            frames = generate_frames(config, [0], RING, 1)
            r = Runner(torch, rsvio_b200, config, frames, local_rank, stream)
            sync = torch.cuda.synchronize
            r.run("device", 0, warmup)
            d, ms, _, _ = r.timed("device", warmup, steps, sync)
            stage_ms, _ = r.stages("device", warmup + steps, 12)
            r.tr.reset()
            r.run("host", 0, warmup)
            dh, msh, wallh, recs = r.timed("host", warmup, steps, sync)
            msh = max(msh, wallh * 1e3)
            alive = int(r.tr.result_view()[1].sum())
            r.close()
            ct, cf, cdt = cpu_run(config, frames, 1, 8, 3, cores)
            alg = 2 * (p["w"] * p["h"] + rsvio_b200.pyramid_bytes(p["w"], p["h"], p["levels"]))
            out.append({
                "config": config, "shape": f"{p['w']}x{p['h']} L{p['levels']} grid {p['grid']}", "streams": 1,
                "value": d["tracks"] / (ms * 1e-3), "unit": "tracks/s", "ms_per_frame": ms / steps,
                "stereo_frames_per_sec": steps / (ms * 1e-3), "tracks_alive": alive,
                "e2e": {"value": dh["tracks"] / (msh * 1e-3), "ms_per_frame": msh / steps,
                        "h2d_bytes_per_step": 2 * p["w"] * p["h"], "d2h_bytes_per_step": recs * 16 / steps},
                "stage_ms": stage_ms,
                "pyramid_latency": {"ms": stage_ms["pyramid"], "alg_bytes": alg,
                                    "note": "2 images: latency-bound, not a bandwidth number"},
                "cpu_baseline": {"value": ct / cdt, "unit": "tracks/s", "cores": cores, "kind": "port",
                                 "stereo_frames_per_sec": cf / cdt, "sample": "1 stream x 8 frames after 3 warm-up frames"},
            })
        except Exception as ex:  # the headline line does not depend on these
            out.append({"config": config, "error": repr(ex)})
    return out


def crate_b_result(rsvio_b200, frames=40, warm=6):
    """Row f-2 (`bench.py --config B1` in VERDICT's words; carried in per_config): the reference's second tracker
    (`feature_tracker` crate, mono, f32 pyramid, bicubic LK, Shi-Tomasi) with its config.yaml settings on the EuRoC-size
    left images of the C1 sequence, through the synchronous rsvio_b_tracker_process_frame (host image in, features on
    the host out).  `value`: device time of a frame (events around the kernels), `e2e`: wall clock of the call."""
    try:
        p = CONFIGS["C1"]
        ring = generate_frames("C1", [0], RING, 1)
        imgs = [(ring[t, 0, 0].astype(np.float32) / np.float32(255.0)) for t in range(RING)]
        try:  # pinned host images: the handle DMA-reads them in place (pageable ones go through its staging copy)
            import torch
            keep = [torch.from_numpy(a).pin_memory() for a in imgs]
            imgs = [k.numpy() for k in keep]
        except Exception:
            pass
        tr = rsvio_b200.FeatureTrackerB(p["w"], p["h"], detection_threshold=0.02)
        dev_ms, wall, tracks, n_feat = [], [], 0, []
        for i in range(warm + frames):
            img = imgs[ring_index(i)]
            t0 = time.perf_counter()
            feats = tr.process_frame(img, i)
            t1 = time.perf_counter()
            if i >= warm:
                wall.append(t1 - t0)
                dev_ms.append(tr.last_frame_ms())
                tracks += len(feats) - tr.n_new          # features of the previous frame that were tracked + kept
                n_feat.append(len(feats))
        prev_total = sum(n_feat[:-1]) + n_feat[0]        # every feature of a previous frame is one forward+backward track
        tr.close()
        dsum, wsum = sum(dev_ms) * 1e-3, sum(wall)
        return {"config": "B1", "shape": f"{p['w']}x{p['h']} mono, crate B (feature_tracker/config/config.yaml: 5 levels, "
                                         "blur 2.0, min_dist 15, detection blur 6.0, 25 iterations; threshold 0.02)",
                "streams": 1, "value": prev_total / dsum, "unit": "tracks/s", "ms_per_frame": 1e3 * dsum / frames,
                "frames_per_sec": frames / dsum, "features_per_frame": float(np.mean(n_feat)),
                "tracked_fraction": tracks / max(prev_total, 1),
                "e2e": {"value": prev_total / wsum, "ms_per_frame": 1e3 * wsum / frames, "frames_per_sec": frames / wsum,
                        "h2d_bytes_per_step": 4 * p["w"] * p["h"], "d2h_bytes_per_step": 16 * float(np.mean(n_feat))},
                "cpu_baseline": None,
                "note": "no CPU baseline: the crate B oracle is a pure-Python restatement (seconds per small frame), "
                        "not a timing reference"}
    except Exception as ex:
        return {"config": "B1", "error": repr(ex)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="C5", choices=list(CONFIGS))
    ap.add_argument("--streams-per-gpu", type=int, default=None,
                    help="weak scaling (default): every GPU runs this many streams (C5: 64)")
    ap.add_argument("--total-streams", type=int, default=None,
                    help="strong scaling: this many streams block-partitioned over the GPUs")
    ap.add_argument("--host-memory", default="wc", choices=["wc", "pinned"],
                    help="frame ring: write-combined pinned pages from rsvio_host_alloc_flags (default) or torch.pin_memory")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-pyramid-roofline", action="store_true")
    ap.add_argument("--no-per-config", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.streams_per_gpu is None:
        args.streams_per_gpu = CONFIGS[args.config].get("streams", 1)
    rank, local_rank, world = sharding.rank_info()

    if args.impl == "reference":
        run_reference(args, rank)
        return

    if args.total_streams:   # strong scaling: fixed job, block partition (rsvio_b200.sharding)
        stream_ids = sharding.partition_streams(args.total_streams, world, rank)
        scaling = "strong"
    else:                    # weak scaling: fixed work per GPU
        stream_ids = sharding.weak_streams(args.streams_per_gpu, rank)
        scaling = "weak"
    S = len(stream_ids)
    if S < 1:
        raise SystemExit("bench.py: this rank has no streams")
    args.streams_per_gpu = S
    p = CONFIGS[args.config]
    w, h = p["w"], p["h"]
    numa = bind_to_gpu_numa_node(local_rank) if world > 1 else None
    cores = os.cpu_count() or 1
    # synthetic frames first (fork pool, before CUDA is initialised); distinct seeds per rank and stream
    t_gen = time.perf_counter()
    frames_np = generate_frames(args.config, stream_ids, RING, max(1, min(16, cores // max(world, 1))))
    t_gen = time.perf_counter() - t_gen

    import torch
    import rsvio_b200
    if not torch.cuda.is_available() or rsvio_b200.device_count() < 1:
        raise SystemExit("bench.py: no CUDA device visible -- the product has no CPU path "
                         "(use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def reduce_max(x):
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def reduce_sum(x):
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    stream = torch.cuda.Stream(device=dev)
    sampler = ClockSampler(local_rank)
    sampler.start()
    run = Runner(torch, rsvio_b200, args.config, frames_np, local_rank, stream, args.host_memory)
    img_bytes = w * h

    # ---- device-resident throughput (`value`) ---------------------------------------------------
    run.run("device", 0, args.warmup)
    d, ms, _, _ = run.timed("device", args.warmup, args.steps, barrier, sampler)
    tracks_dev, frames_dev, iters_dev, ms_dev = sharding.aggregate(dist, dev, d["tracks"], S * args.steps,
                                                                   d["iterations"], ms)
    templ_dev = reduce_sum(d["templates"])
    launches = d["launches"]
    n_tracks_alive = int(run.tr.result_view()[1].sum())

    # ---- per-stage device times of that path (profiling mode, device-resident frames) -----------
    prof_steps = max(10, min(50, args.steps))
    sampler.active = True
    stages, stage_launches = run.stages("device", args.warmup + args.steps, prof_steps)
    sampler.active = False

    # ---- end to end: pinned host frames in, host-mapped result out (`e2e`) -----------------------
    run.tr.reset()
    run.run("host", 0, args.warmup)
    dh, ms_e2e_dev, wall_e2e, recs = run.timed("host", args.warmup, args.steps, barrier, sampler)
    ms_e2e_dev = reduce_max(ms_e2e_dev)
    ms_e2e = max(ms_e2e_dev, reduce_max(wall_e2e * 1e3))   # host reads are part of the step
    tracks_e2e, frames_e2e, _, _ = sharding.aggregate(dist, dev, dh["tracks"], S * args.steps, 0, 0.0)
    d2h_per_step = (recs * 16 + 2 * S * 4) / args.steps
    stages_host, _ = run.stages("host", args.warmup + args.steps, min(prof_steps, 20))
    h2d_gbs_rank = 2 * S * img_bytes / (ms_e2e / args.steps * 1e-3) / 1e9

    total_stage_ms = sum(v for k, v in stages.items() if k != "upload")
    pyr_launches = max(1, stage_launches["pyramid"])          # kernels per pyramid build
    pyr_ms = stages["pyramid"]                                 # all of them: the whole pyramid of a step
    alg_bytes_launch = 2 * S * (w * h + rsvio_b200.pyramid_bytes(w, h, p["levels"]))
    peak, peak_src = measured_peaks()
    achieved = alg_bytes_launch / (pyr_ms * 1e-3) / 1e9 if pyr_ms > 0 else 0.0
    ncu_traffic, ncu_traffic_src = ncu_dram_traffic(S == 64 and args.config == "C5")
    roofline = {"kernel": "pyr_fast_kernel (pyramid build of one step: %d launch(es), both cameras, all streams)" % pyr_launches,
                "bound": "hbm", "achieved": achieved,
                "peak": peak, "peak_source": peak_src, "unit": "GB/s", "frac": achieved / peak, "traffic": ncu_traffic,
                "traffic_source": ncu_traffic_src,
                "traffic_over_algorithmic": (ncu_traffic / alg_bytes_launch) if ncu_traffic else None,
                "alg_bytes_per_launch": alg_bytes_launch, "ms_per_launch": pyr_ms, "launches_per_step": pyr_launches,
                "how": "CUDA events on the handle's stream around the pyramid launch of every profiled step "
                       "(device-resident frames, input ring of 8 frames x 46 MB: cold inputs)",
                "share_of_step": pyr_ms / total_stage_ms if total_stage_ms else None}

    # ---- LK: not HBM-bound; against the FP32 issue rate measured in this run ---------------------
    lk_ms = stages["lk_temporal"] + stages["lk_stereo"]
    lk_roofline = None
    try:
        ffma_tf, split_tf = rsvio_b200.microbench_fp32(local_rank)
        step_ms = ms_dev / args.steps
        flops_step = (6.3e3 * templ_dev + 1.8e3 * iters_dev) / args.steps / max(world, 1)
        lk_roofline = {"kernel": "lk_track_fb_kernel (temporal + stereo launch of a step)", "bound": "fp32 issue",
                       "achieved": flops_step / (lk_ms * 1e-3) / 1e12 if lk_ms > 0 else None, "unit": "TFLOP/s",
                       "peak_ffma": ffma_tf, "peak_fmul_fadd": split_tf,
                       "frac_of_ffma": flops_step / (lk_ms * 1e-3) / 1e12 / ffma_tf if lk_ms > 0 and ffma_tf else None,
                       "algorithmic_flops": "6.3 kflop per template + 1.8 kflop per iteration (SURVEY 8d), counted by the kernel",
                       "ms_per_step": lk_ms, "share_of_step": lk_ms / total_stage_ms if total_stage_ms else None,
                       "tracks_per_step": tracks_dev / args.steps / max(world, 1),
                       "note": "warp-instruction counts per track: profiles/ (ncu); most issue slots are addressing, "
                               "u8 gathers + conversions, predicates and shuffles, not flops"}
        del step_ms
    except Exception as ex:
        lk_roofline = {"error": repr(ex)}

    # ---- pyramid kernels alone on batches far larger than L2 (>= 512 MB per launch) --------
    extra_roofline = None
    if not args.no_pyramid_roofline and rank == 0 and world == 1:
        import ctypes as C
        extra_roofline = []
        L = rsvio_b200.load_library()
        for cw, ch, cl, n_img, label in ((w, h, p["levels"], 640, "bench shape"), (1920, 1200, 5, 96, "C4 shape"),
                                         (752, 480, 6, 128, "shipped EuRoC L6, one step's batch"),
                                         (800, 400, 6, 128, "C3 L6")):
            try:
                src = torch.randint(0, 256, (n_img, ch, cw), dtype=torch.uint8, device=dev)
                pb = L.rsvio_pyramid_device_bytes(cw, ch, cl)
                dst = torch.empty(n_img * pb, dtype=torch.uint8, device=dev)
                msv = C.c_float(0)
                for rep in (3, 10):
                    st = L.rsvio_build_pyramid_u8_device(local_rank, C.c_void_p(src.data_ptr()), n_img, cw, ch, cw, cw * ch,
                                                         cl, C.c_void_p(dst.data_ptr()), rep, C.byref(msv))
                    if st != 0:
                        raise RuntimeError(f"status {st}")
                ab = n_img * (cw * ch + rsvio_b200.pyramid_bytes(cw, ch, cl))
                extra_roofline.append({"workload": f"{n_img} x {cw}x{ch} {cl}-level pyramids per launch set "
                                                   f"({ab / 1e6:.0f} MB algorithmic, {label})",
                                       "achieved": ab / (msv.value * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                                       "frac": ab / (msv.value * 1e-3) / 1e9 / peak, "ms_per_launch_set": msv.value})
                del src, dst
            except Exception as ex:  # the headline number does not depend on this microbenchmark
                extra_roofline.append({"error": str(ex)})

    # ---- latency of ONE stream through the synchronous reference-facing call (the drop-in use case) ----
    single = None
    if rank == 0:
        import ctypes as C
        single = {}
        for name, lv, grid in (("L3", p["levels"], p["grid"]), ("shipped_L6", 6, 50)):
            try:
                tr1 = rsvio_b200.StereoPatchTracker(grid, p["max_it"], p["thr"], width=w, height=h, levels=lv,
                                                    device=local_rank)
                base = run.hostbuf.ptr if run.host_kind == "wc" else run.hostbuf.data_ptr()
                lat = []
                for i in range(60):
                    k = ring_index(i)
                    lp = base + (k * 2 + 0) * S * img_bytes
                    rp = base + (k * 2 + 1) * S * img_bytes
                    t0 = time.perf_counter()
                    st = tr1._L.rsvio_tracker_process_stereo(tr1._h, C.c_void_p(lp), C.c_void_p(rp), w)
                    lat.append(time.perf_counter() - t0)
                    if st != 0:
                        raise RuntimeError(f"status {st}")
                lat = np.array(lat[10:]) * 1e3
                single[name] = {"ms_per_frame_median": float(np.median(lat)), "ms_per_frame_p95": float(np.percentile(lat, 95)),
                                "stereo_frames_per_sec": float(1e3 / np.median(lat)),
                                "tracks_alive": tr1.num_tracks(0) + tr1.num_tracks(1),
                                "what": f"1 stream, {lv} levels, grid {grid}: rsvio_tracker_process_stereo per frame "
                                        "(upload + kernels + result), host wall clock"}
                tr1.close()
            except Exception as ex:
                single[name] = {"error": str(ex)}

    run.close()

    # ---- the other BASELINE configs, one stream each (short) ---------------------------------------
    per_config = None
    if rank == 0 and world == 1 and not args.no_per_config:
        per_config = per_config_results(torch, rsvio_b200, local_rank, stream, cores,
                                        ("C1", "C2", "C3", "C4", "EUROC_SHIPPED"))
        per_config.append(crate_b_result(rsvio_b200))

    sampler.stop_flag = True
    clocks = sampler.summary()

    # ---- CPU baseline (oracle port) on a bounded sample, rank 0 at N = 1 only -------------------------
    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        ns = min(S, 2 * cores if cores < 32 else 64)
        # calibrate on a few frames, then size the sample for roughly 12 s of CPU work on this box
        _, _, dt0 = cpu_run(args.config, frames_np, ns, 8, 2, cores)
        nf = int(max(20, min(4000, 8 * 12.0 / max(dt0, 1e-3))))
        tracks, nframes, dt = cpu_run(args.config, frames_np, ns, nf, 2, cores)
        cpu_baseline = {"value": tracks / dt, "unit": "tracks/s", "cores": min(cores, ns), "kind": "port",
                        "stereo_frames_per_sec": nframes / dt,
                        "sample": f"{ns} streams x {nf} frames after 2 warm-up frames, oracle port of the reference "
                                  f"tracker (C++ -O3, f32, no FMA), one stream per thread, {dt:.1f} s"}

    if rank == 0:
        value = tracks_dev / (ms_dev * 1e-3)
        line = {
            "metric": "feature_tracks_per_sec", "value": value, "unit": "tracks/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_dev / args.steps,
            "higher_is_better": True, "scaling": scaling, "vs_baseline": None, "dtype": "f32 (u8 pyramid)",
            "data": "synthetic",
            "config": workload_config(args, S, world),
            "stereo_frames_per_sec": frames_dev / (ms_dev * 1e-3),
            "lk_iterations_per_sec": iters_dev / (ms_dev * 1e-3),
            "lk_gflops": (6.3e3 * templ_dev + 1.8e3 * iters_dev) / (ms_dev * 1e-3) / 1e9,
            "tracks_alive_per_gpu": n_tracks_alive,
            "e2e": {"value": tracks_e2e / (ms_e2e * 1e-3), "unit": "tracks/s",
                    "stereo_frames_per_sec": frames_e2e / (ms_e2e * 1e-3),
                    "ms_per_step": ms_e2e / args.steps, "ms_per_step_device_events": ms_e2e_dev / args.steps,
                    "h2d_bytes_per_step": 2 * S * img_bytes, "d2h_bytes_per_step": d2h_per_step,
                    "h2d_gbs_per_gpu": h2d_gbs_rank, "h2d_gbs_all_gpus": h2d_gbs_rank * world,
                    "host_memory": args.host_memory, "result_read": "every record of every step (x and y summed on the host)"},
            "gpu_launches": int(launches),
            "roofline": roofline,
            "roofline_lk": lk_roofline,
            "roofline_pyramid_batch": extra_roofline,
            "stage_ms_per_step": stages,
            "stage_share": {k: (v / total_stage_ms if (total_stage_ms and k != "upload") else None) for k, v in stages.items()},
            "stage_ms_per_step_host_frames": stages_host,
            "per_config": per_config,
            "cpu_baseline": cpu_baseline,
            "single_stream_latency": single,
            "clocks": clocks,
            "datagen_s": t_gen,
            "numa_bound_cpus": numa,
        }
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
