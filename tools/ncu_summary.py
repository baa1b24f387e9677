#!/usr/bin/env python
"""Condense an .ncu-rep (ncu --set full) into the few numbers the design cares about.

usage: ncu_summary.py <report.ncu-rep> <reactions-in-launch> [libsonics_b200.so kernel-substring]
"""
import csv
import io
import json
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "smsp__thread_inst_executed.sum",
    "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tc.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "sm__cycles_elapsed.max", "smsp__cycles_active.avg",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
]


def main():
    rep = sys.argv[1]
    reactions = float(sys.argv[2])
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    d = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
    print("kernel:", d.get("Kernel Name", ("?",))[0])
    res = {}
    for k in KEYS:
        if k in d:
            print("%-82s %16s %s" % (k, d[k][0], d[k][1]))
            try:
                res[k] = float(d[k][0].replace(",", ""))
            except ValueError:
                pass
    wi = res.get("smsp__inst_executed.sum")
    if wi:
        unit_scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        dram = sum(res.get(k, 0) * unit_scale.get(d[k][1], 1) for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
        summary = {"reactions_in_launch": reactions, "warp_inst_per_reaction": wi / reactions,
                   "thread_inst_per_reaction": res.get("smsp__thread_inst_executed.sum", 0) / reactions,
                   "avg_active_threads": res.get("smsp__thread_inst_executed_per_inst_executed.ratio"),
                   "issue_active_pct": res.get("smsp__issue_active.avg.pct_of_peak_sustained_active"),
                   "dram_bytes_per_launch": dram, "source": rep.split("/")[-1]}
        print("\nsummary:", json.dumps(summary))
    if len(sys.argv) > 4:
        print()
        sys.stdout.flush()
        subprocess.call([sys.executable, __file__.replace("ncu_summary.py", "ncu_lines.py"), rep, sys.argv[3], sys.argv[4], "30"])


if __name__ == "__main__":
    main()
