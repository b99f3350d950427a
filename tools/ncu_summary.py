"""Summaries of ncu outputs for profiles/ (read here, on the CPU box).
  python tools/ncu_summary.py launches gpurun_out/X_launches.csv   > profiles/...
  python tools/ncu_summary.py kernel   gpurun_out/X.ncu-rep        > profiles/..."""
import collections
import csv
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size",
    "launch__registers_per_thread", "launch__occupancy_limit_registers",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "sm__inst_executed_pipe_alu.sum.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.sum.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.sum.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.sum.pct_of_peak_sustained_active",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.per_second",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
    "lts__t_bytes.sum", "l1tex__t_bytes.sum",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
]


def launches(path):
    rows = list(csv.reader(l for l in open(path) if l.startswith('"')))
    h = rows[0]
    ki, vi, mi = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Name")
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows[1:]:
        if r[mi] != "gpu__time_duration.sum":
            continue
        n = r[ki].split("(")[0]
        agg[n][0] += 1
        agg[n][1] += float(r[vi].replace(",", ""))
    tot = sum(v[1] for v in agg.values())
    print(f"# ncu --metrics gpu__time_duration.sum --clock-control none ; source {path}")
    print(f"# {'kernel':30s} {'launches':>8s} {'total_ms':>10s} {'avg_us':>10s} {'share':>6s}")
    for n, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
        print(f"  {n:30s} {c:8d} {t / 1e6:10.3f} {t / c / 1e3:10.1f} {t / tot:6.3f}")


def kernel(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    h, units = rows[0], rows[1]
    print(f"# ncu --set full --clock-control none ; source {path}")
    for r in rows[2:]:
        print(f"## kernel: {r[h.index('Kernel Name')]}  grid={r[h.index('Grid Size')]} block={r[h.index('Block Size')]}")
        for k in KEYS:
            if k in h:
                i = h.index(k)
                print(f"  {k:85s} {r[i]:>16s} {units[i]}")


def traffic(path):
    """profiles/r2_traffic.json: dram bytes (read + write) per launch of every captured kernel,
    averaged over the captured launches - what bench.py reports as roofline.traffic"""
    import json

    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv", "--print-units", "base"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    h = rows[0]
    ki, ri, wi, ti = (h.index("Kernel Name"), h.index("dram__bytes_read.sum"), h.index("dram__bytes_write.sum"),
                      h.index("gpu__time_duration.sum"))
    agg = collections.defaultdict(lambda: [0, 0.0, 0.0])
    for r in rows[2:]:
        n = r[ki].split("(")[0]
        agg[n][0] += 1
        agg[n][1] += float(r[ri].replace(",", "")) + float(r[wi].replace(",", ""))
        agg[n][2] += float(r[ti].replace(",", ""))
    res = {"source": f"ncu --set full --clock-control none, {path.split('/')[-1]} (summary: profiles/)"}
    for n, (c, b, t) in agg.items():
        res[n] = {"dram_bytes": int(b / c), "launches_captured": c, "avg_us_under_ncu": round(t / c / 1e3, 1)}
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    {"launches": launches, "kernel": kernel, "traffic": traffic}[sys.argv[1]](sys.argv[2])
