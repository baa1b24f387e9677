#!/usr/bin/env python
"""Aggregate an ncu capture of k_pcr by code region (phase of the state machine / sampler piece).

usage: ncu_regions.py <report.ncu-rep> <libsonics_b200.so> [kernel-substring=k_pcr]
"""
import csv
import io
import os
import subprocess
import sys
from collections import defaultdict

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ncu_lines  # noqa: E402

CSRC = os.path.join(HERE, "..", "sonics_b200", "csrc")


def marks(path, pairs):
    src = open(path).read().split("\n")

    def find(s):
        for i, x in enumerate(src):
            if s in x:
                return i + 1
        raise KeyError(s)
    out = {}
    for name, a, b in pairs:
        out[name] = (find(a), (find(b) - 1) if b else len(src))
    return out


SIM = marks(os.path.join(CSRC, "sim_kernel.cuh"), [
    ("boundary_phase", "// ================= BOUNDARY (deferred)", "// ================= ADVANCE"),
    ("adv1_consume", "// ---- (1) consume the finished draw", "// ---- (2) walk to the next draw"),
    ("adv2_walk", "// ---- (2) walk to the next draw", "// ---- (3) classify the draw"),
    ("adv3_classify", "// ---- (3) classify the draw", "// ================= INV (deferred)"),
    ("inv_trial_phase", "// ================= INV (deferred)", "// ================= deferred Stirling test"),
    ("full_phase", "// ================= deferred Stirling test", "// K1b: one lane per simulation"),
])
SMP = marks(os.path.join(CSRC, "samplers.cuh"), [
    ("stirling_tail", "float stirling_tail", "float floor_small"),
    ("inv_setup", "void inv_setup", "void inv_start"),
    ("inv_run", "void inv_start", "// ---- BTRS"),
    ("btrs_setup", "void btrs_setup", "// One rejection trial"),
    ("btrs_trial", "// One rejection trial", "// The exact test"),
    ("btrs_full", "// The exact test", "// Binomial(n, p), any p"),
    ("poisson", "// ---- Poisson", None),
])


LS = marks(os.path.join(CSRC, "sim_lockstep.cuh"), [
    ("ls_sort_keys", "__global__ void k_sort_keys", "// ---- lockstep samplers"),
    ("ls_binomial", "int binomial_ls(", "int poisson_ls("),
    ("ls_poisson", "int poisson_ls(", "// Shared memory: [rcp table]"),
    ("ls_batch_setup", "// next batch of 32 units", "// ---- simulate (:397-480)"),
    ("ls_capture", "// ---- simulate (:397-480)", "// ---- amplification sweep of this cycle"),
    ("ls_sweep", "// ---- amplification sweep of this cycle", "// ---- lnL (+ readout) of the final pool"),
    ("ls_finish", "// ---- lnL (+ readout) of the final pool", None),
])


def region(loc):
    if not loc:
        return "none"
    f, l = loc
    tab = SIM if f == "sim_kernel.cuh" else SMP if f == "samplers.cuh" else LS if f == "sim_lockstep.cuh" else None
    if tab is None:
        return f
    for name, (a, b) in tab.items():
        if a <= l <= b:
            return name
    return f + ":other"


def main():
    rep, so = sys.argv[1:3]
    ksub = sys.argv[3] if len(sys.argv) > 3 else "k_pcr"
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[hi]
    ci, ti = hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed")
    data = [r for r in rows[hi + 1:] if len(r) > ti]
    table = ncu_lines.line_table(so, ksub)
    assert len(table) == len(data), (len(table), len(data))
    agg = defaultdict(lambda: [0, 0, 0])
    tot = 0
    for (off, loc, text), r in zip(table, data):
        w, t = int(r[ci]), int(r[ti])
        a = agg[region(loc)]
        a[0] += w
        a[1] += t
        a[2] += 1
        tot += w
    print("%-20s %10s %8s %6s %12s" % ("region", "warp-inst%", "avg-act", "SASS", "thread-inst%"))
    tt = sum(v[1] for v in agg.values())
    for r, (w, t, n) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
        print("%-20s %10.2f %8.2f %6d %12.2f" % (r, 100.0 * w / tot, t / max(1, w), n, 100.0 * t / tt))


if __name__ == "__main__":
    main()
