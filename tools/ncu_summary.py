"""Summarise ncu artefacts into small text files for profiles/.
  python tools/ncu_summary.py launches <launches.csv> <out.txt>     per-kernel launch count / avg / share
  python tools/ncu_summary.py kernel   <report.ncu-rep> <out.txt>   key metrics of the longest launch in the report"""
import collections
import csv
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
]


def launches(src, dst):
    rows = list(csv.reader(open(src)))
    hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    H, data = rows[hdr], rows[hdr + 1:]
    ki, vi, gi = H.index("Kernel Name"), H.index("Metric Value"), H.index("Grid Size")
    agg = collections.defaultdict(list)
    n_lk, after_detect = 0, False
    for r in data:
        if len(r) > vi:
            name = r[ki].replace("unnamed>::", "").replace("void ", "").replace("rsvio::<", "").split("(")[0]
            if "pyr_" in name:
                n_lk, after_detect = 0, False
            if "detect_list" in name:
                after_detect = True
            if "lk_track_fb" in name:
                # launches of a frame in host order: temporal (split on batched handles: survivors of the previous
                # compaction, then the newly merged points, then the helper that joins the first), stereo after detection
                n_lk += 1
                if after_detect:
                    name += " [stereo, new points]"
                elif n_lk == 1:
                    name += " [temporal]"
                elif n_lk == 2:
                    name += " [temporal, points merged last frame]"
                else:
                    name += " [temporal, helper: queue already empty under ncu]"
            agg[f"{name} grid={r[gi]}"].append(float(r[vi].replace(",", "")))
    tot = sum(sum(v) for v in agg.values())
    with open(dst, "w") as f:
        f.write("# ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised: compare shares);\n"
                "# rows are (kernel, grid): the bench's 64-stream handle and its 1-stream latency probe launch the same kernels\n")
        f.write(f"{'kernel':70s} {'n':>5s} {'avg_us':>10s} {'share':>7s}\n")
        for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
            f.write(f"{k[:70]:70s} {len(v):5d} {sum(v) / len(v) / 1e3:10.1f} {sum(v) / tot:7.3f}\n")


def kernel(src, dst):
    out = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    H, U = rows[0], rows[1]
    ti = H.index("gpu__time_duration.sum")
    D = max(rows[2:], key=lambda r: float(r[ti].replace(",", "")) if len(r) > ti and r[ti] else 0.0)  # the longest launch captured
    with open(dst, "w") as f:
        f.write(f"# {src}: ncu --set full --clock-control none, kernel: {D[H.index('Kernel Name')]}\n")
        for k in KEYS:
            if k in H:
                i = H.index(k)
                f.write(f"{k:75s} {D[i]:>20s} {U[i]}\n")
        for i, h in enumerate(H):
            if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio"):
                if float(D[i] or 0) >= 0.1:
                    f.write(f"{h:75s} {D[i]:>20s}\n")


if __name__ == "__main__":
    {"launches": launches, "kernel": kernel}[sys.argv[1]](sys.argv[2], sys.argv[3])
