#!/usr/bin/env python
"""Summarise ncu reports (.ncu-rep, `--set full`) and a launch list (csv) into profiles/<tag>_ncu_summary.txt.
Usage: python tools/ncu_summary.py <tag>
Reads gpurun_out/launches_<tag>.csv, gpurun_out/prof_<tag>.ncu-rep and (if present) gpurun_out/prof_fp32_<tag>.ncu-rep."""
import collections
import csv
import os
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "sm__cycles_elapsed.avg", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
        "TPC.TriageCompute.sm__pipe_tensor_subpipe_imma_cycles_active_realtime.avg",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__occupancy_limit_registers",
        "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio"]


def launch_list(path, out):
    rows = list(csv.reader(open(path)))
    h = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    H = rows[h]
    ki, vi = H.index("Kernel Name"), H.index("Metric Value")
    agg = collections.OrderedDict()
    for r in rows[h + 1:]:
        if len(r) > vi and r[vi].replace(",", "").replace(".", "").isdigit():
            a = agg.setdefault(r[ki][:90], [0, 0.0])
            a[0] += 1
            a[1] += float(r[vi].replace(",", ""))
    tot = sum(v[1] for v in agg.values())
    out.write("== launch list (%s): ncu --metrics gpu__time_duration.sum --clock-control none -c 400 python bench.py --steps 4 --warmup 3 ==\n" % path)
    for k, v in sorted(agg.items(), key=lambda x: -x[1][1]):
        out.write("%-92s n=%4d  total %10.1f us  avg %9.1f us  share %5.1f%%\n" % (k, v[0], v[1] / 1e3, v[1] / 1e3 / v[0], 100 * v[1] / tot))


def full_report(rep, out):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    H = rows[0]
    idx = {x: i for i, x in enumerate(H)}
    out.write("\n== ncu --set full --clock-control none --import-source on (%s) ==\n" % rep)
    for r in rows[2:]:
        out.write("kernel: %s\n" % r[idx["Kernel Name"]][:140])
        for w in WANT:
            if w in idx:
                out.write("  %-96s %s %s\n" % (w, r[idx[w]], rows[1][idx[w]]))


def main():
    tag = sys.argv[1]
    out = open("profiles/%s_ncu_summary.txt" % tag, "w")
    try:
        launch_list("gpurun_out/launches_%s.csv" % tag, out)
    except Exception as e:  # noqa
        out.write("launch list unavailable: %s\n" % e)
    for rep in ("gpurun_out/prof_%s.ncu-rep" % tag, "gpurun_out/prof_fp32_%s.ncu-rep" % tag):
        if os.path.exists(rep):
            try:
                full_report(rep, out)
            except Exception as e:  # noqa
                out.write("ncu report %s unavailable: %s\n" % (rep, e))
    out.close()
    print(open("profiles/%s_ncu_summary.txt" % tag).read())


if __name__ == "__main__":
    main()
