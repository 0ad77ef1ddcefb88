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
    GIL), frames [0, warm) untimed.  Returns (tracks, stereo_frames, seconds)."""
    from concurrent.futures import ThreadPoolExecutor
    from oracle import oracle as orc
    orc.build()
    p = CONFIGS[config]
    trackers = [orc.StereoTracker(p["levels"], p["grid"], p["max_it"], p["thr"], nthreads=1) for _ in range(n_streams)]

    def advance(s, t0, t1):
        for t in range(t0, t1):
            k = ring_index(t, frames.shape[0])
            trackers[s].process_frame(frames[k, 0, s], frames[k, 1, s])

    with ThreadPoolExecutor(max_workers=threads) as ex:
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
            "parallelism": f"{n_gpus} x independent replicas (streams sharded, no collective)"}


# ------------------------------------------------------------------------------------------------
def ncu_dram_traffic(same_workload):
    """dram__bytes_read.sum + dram__bytes_write.sum of the pyramid kernel, per launch, from the committed
    `ncu --set full` capture of this same workload (profiles/r1_pyr_fast_kernel.txt); None for other workloads."""
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", "r1_pyr_fast_kernel.txt")
    if not same_workload or not os.path.exists(path):
        return None, None
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    tot = 0.0
    for line in open(path):
        f = line.split()
        if len(f) >= 3 and f[0] in ("dram__bytes_read.sum", "dram__bytes_write.sum") and f[2] in scale:
            tot += float(f[1]) * scale[f[2]]
    return (tot, "profiles/r1_pyr_fast_kernel.txt (ncu --set full, cold L2, one launch of 128 images)") if tot else (None, None)


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
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-pyramid-roofline", action="store_true")
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

    host = torch.from_numpy(frames_np).pin_memory()                  # [ring][cam][S][h][w]
    devbuf = host.to(dev)
    img_bytes = w * h
    stream = torch.cuda.Stream(device=dev)
    sampler = ClockSampler(local_rank)
    sampler.start()

    import ctypes as C
    host_ptrs = []
    for k in range(RING):
        lp = (C.c_void_p * S)(*[host[k, 0, s].data_ptr() for s in range(S)])
        rp = (C.c_void_p * S)(*[host[k, 1, s].data_ptr() for s in range(S)])
        host_ptrs.append((lp, rp))

    with torch.cuda.stream(stream):
        tr = rsvio_b200.StereoPatchTracker(p["grid"], p["max_it"], p["thr"], width=w, height=h, levels=p["levels"],
                                           n_streams=S, device=local_rank, cuda_stream=stream.cuda_stream)

        def submit_dev(i):
            k = ring_index(i)
            tr.submit_device(devbuf[k, 0].data_ptr(), devbuf[k, 1].data_ptr(), w, img_bytes)

        def run_device(first, n):
            """n pipelined steps (<= 2 in flight) from device-resident frames"""
            submit_dev(first)
            for i in range(first + 1, first + n):
                submit_dev(i)
                tr.sync()
            tr.sync()

        # ---- device-resident throughput --------------------------------------------------------
        run_device(0, args.warmup)
        barrier()
        c0 = tr.counters()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        sampler.active = True
        e0.record(stream)
        run_device(args.warmup, args.steps)
        e1.record(stream)
        e1.synchronize()
        sampler.active = False
        barrier()
        c1 = tr.counters()
        tracks_dev, frames_dev, iters_dev, ms_dev = sharding.aggregate(
            dist, dev, c1["tracks"] - c0["tracks"], S * args.steps, c1["iterations"] - c0["iterations"],
            e0.elapsed_time(e1))
        templ_dev = reduce_sum(c1["templates"] - c0["templates"])
        launches = c1["launches"] - c0["launches"]
        n_tracks_alive = int(tr.result_view()[1].sum())

        # ---- end to end: pinned host frames in, host-mapped result out ---------------------------
        tr.reset()

        def submit_host(i):
            lp, rp = host_ptrs[ring_index(i)]
            tr.submit_batch_ptr(lp, rp, w)

        consumed = [0, 0.0]

        def consume():
            rec, cnt = tr.result_view()              # zero-copy view of host-mapped pinned memory
            n = int(cnt.sum())
            consumed[0] += n
            # read the step's result on the host: x of the first track of every group
            consumed[1] += float(rec["x"][:, 0].sum())
            return n

        def run_host(first, n):
            d2h = 0
            submit_host(first)
            for i in range(first + 1, first + n):
                submit_host(i)
                tr.sync()
                d2h += consume()
            tr.sync()
            d2h += consume()
            return d2h

        run_host(0, args.warmup)
        barrier()
        c2 = tr.counters()
        barrier()
        sampler.active = True
        e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e2.record(stream)
        recs = run_host(args.warmup, args.steps)
        e3.record(stream)
        e3.synchronize()
        wall_e2e = time.perf_counter() - t0
        sampler.active = False
        barrier()
        ms_e2e_dev = reduce_max(e2.elapsed_time(e3))
        ms_e2e = max(ms_e2e_dev, reduce_max(wall_e2e * 1e3))   # host reads are part of the step
        c3 = tr.counters()
        tracks_e2e, frames_e2e, _, _ = sharding.aggregate(dist, dev, c3["tracks"] - c2["tracks"], S * args.steps, 0, 0.0)
        d2h_per_step = (recs * 16 + 2 * S * 4) / args.steps

        # ---- per-stage device times (CUDA events inside the library, same stream) ----------------
        prof_steps = max(10, min(50, args.steps))
        tr.set_profiling(True)
        tr.stage_times(reset=True)
        sampler.active = True
        run_host(args.warmup + args.steps, prof_steps)
        sampler.active = False
        stages = tr.stage_times(reset=True)
        tr.set_profiling(False)

    total_stage_ms = sum(v[0] for k, v in stages.items() if k != "upload")
    pyr_launches = max(1, stages["pyramid"][1] // prof_steps)     # kernels per pyramid build
    pyr_ms = stages["pyramid"][0] / prof_steps                    # all of them: the whole pyramid of a step
    alg_bytes_launch = 2 * S * (w * h + rsvio_b200.pyramid_bytes(w, h, p["levels"]))
    peak, peak_src = measured_peaks()
    achieved = alg_bytes_launch / (pyr_ms * 1e-3) / 1e9 if pyr_ms > 0 else 0.0
    ncu_traffic, ncu_traffic_src = ncu_dram_traffic(S == 64 and args.config == "C5")
    roofline = {"kernel": "pyr_fast_kernel (pyramid build of one step: %d launch(es), both cameras, all streams)" % pyr_launches,
                "bound": "hbm", "achieved": achieved,
                "peak": peak, "peak_source": peak_src, "unit": "GB/s", "frac": achieved / peak, "traffic": ncu_traffic,
                "traffic_source": ncu_traffic_src,
                "alg_bytes_per_launch": alg_bytes_launch, "ms_per_launch": pyr_ms, "launches_per_step": pyr_launches,
                "share_of_step": stages["pyramid"][0] / total_stage_ms if total_stage_ms else None}
    stage_share = {k: (v[0] / prof_steps, v[0] / total_stage_ms if (total_stage_ms and k != "upload") else None)
                   for k, v in stages.items()}

    # ---- pyramid kernels alone on batches far larger than L2 (>= 512 MB per launch) --------
    extra_roofline = None
    if not args.no_pyramid_roofline and rank == 0 and world == 1:
        extra_roofline = []
        L = rsvio_b200.load_library()
        for cw, ch, cl, n_img, label in ((w, h, p["levels"], 640, "bench shape"), (1920, 1200, 5, 96, "C4 shape")):
            try:
                src = torch.randint(0, 256, (n_img, ch, cw), dtype=torch.uint8, device=dev)
                pb = L.rsvio_pyramid_device_bytes(cw, ch, cl)
                dst = torch.empty(n_img * pb, dtype=torch.uint8, device=dev)
                ms = C.c_float(0)
                for rep in (3, 10):
                    st = L.rsvio_build_pyramid_u8_device(local_rank, C.c_void_p(src.data_ptr()), n_img, cw, ch, cw, cw * ch,
                                                         cl, C.c_void_p(dst.data_ptr()), rep, C.byref(ms))
                    if st != 0:
                        raise RuntimeError(f"status {st}")
                ab = n_img * (cw * ch + rsvio_b200.pyramid_bytes(cw, ch, cl))
                extra_roofline.append({"workload": f"{n_img} x {cw}x{ch} {cl}-level pyramids per launch set "
                                                   f"({ab / 1e6:.0f} MB algorithmic, {label})",
                                       "achieved": ab / (ms.value * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                                       "frac": ab / (ms.value * 1e-3) / 1e9 / peak, "ms_per_launch_set": ms.value})
                del src, dst
            except Exception as ex:  # the headline number does not depend on this microbenchmark
                extra_roofline.append({"error": str(ex)})

    # ---- latency of ONE stream through the synchronous reference-facing call (the drop-in use case) ----
    single = None
    if rank == 0:
        try:
            tr1 = rsvio_b200.StereoPatchTracker(p["grid"], p["max_it"], p["thr"], width=w, height=h, levels=p["levels"],
                                                device=local_rank)
            lat = []
            for i in range(60):
                k = ring_index(i)
                lp, rp = host[k, 0, 0], host[k, 1, 0]
                t0 = time.perf_counter()
                rsvio_b200.binding._check(tr1._L.rsvio_tracker_process_stereo(tr1._h, C.c_void_p(lp.data_ptr()),
                                                                                C.c_void_p(rp.data_ptr()), w), tr1._h)
                lat.append(time.perf_counter() - t0)
            lat = np.array(lat[10:]) * 1e3
            single = {"ms_per_frame_median": float(np.median(lat)), "ms_per_frame_p95": float(np.percentile(lat, 95)),
                      "stereo_frames_per_sec": float(1e3 / np.median(lat)), "tracks_alive": tr1.num_tracks(0) + tr1.num_tracks(1),
                      "what": "1 stream, rsvio_tracker_process_stereo per frame (upload + 9 kernels + result), host wall clock"}
            tr1.close()
        except Exception as ex:
            single = {"error": str(ex)}

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
                    "h2d_bytes_per_step": 2 * S * img_bytes, "d2h_bytes_per_step": d2h_per_step},
            "gpu_launches": int(launches),
            "roofline": roofline,
            "roofline_pyramid_batch": extra_roofline,
            "stage_ms_per_step": {k: v[0] for k, v in stage_share.items()},
            "stage_share": {k: v[1] for k, v in stage_share.items()},
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
