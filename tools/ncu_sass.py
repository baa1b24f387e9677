#!/usr/bin/env python
"""SASS of a kernel in address order with the executed counts of an ncu capture (hot instructions only).

usage: ncu_sass.py <report.ncu-rep> <libsonics_b200.so> <kernel-substring> [min_share_percent]"""
import csv
import io
import subprocess
import sys

from ncu_lines import line_table


def main():
    rep, so, ksub = sys.argv[1:4]
    thr = float(sys.argv[4]) if len(sys.argv) > 4 else 0.05
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[hdr_i]
    ci, ti, si = hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed"), hdr.index("# Samples")
    data = [r for r in rows[hdr_i + 1:] if len(r) > ti]
    table = line_table(so, ksub)
    assert len(table) == len(data)
    tot = sum(int(r[ci]) for r in data)
    tots = sum(int(r[si]) for r in data)
    for (off, loc, text), r in zip(table, data):
        w, t, s = int(r[ci]), int(r[ti]), int(r[si])
        if 100.0 * w / tot < thr:
            continue
        print("%05x %6.3f%% act %5.1f smp %5.2f%%  %-24s %s" % (off, 100.0 * w / tot, t / max(1, w), 100.0 * s / tots,
              "%s:%d" % loc if loc else "-", text[:90]))


if __name__ == "__main__":
    main()
