#!/usr/bin/env python
"""Attribute an ncu capture's per-SASS-instruction counters to source lines.

usage: ncu_lines.py <report.ncu-rep> <libsonics_b200.so> <kernel-substring> [top_n]

Joins `ncu --page source --csv` (per-instruction executed counts, in address order)
with `nvdisasm -g` line info of the same cubin, and prints warp-instructions,
thread-instructions and SIMT efficiency per (file, line) and per file.
"""
import csv
import io
import os
import re
import subprocess
import sys
import tempfile
from collections import defaultdict


def line_table(so, kernel_sub):
    tmp = tempfile.mkdtemp()
    subprocess.check_call(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=tmp, stdout=subprocess.DEVNULL)
    cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
    dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
    table = []
    cur = None
    active = False
    for ln in dis.splitlines():
        if ln.startswith(".text."):
            active = kernel_sub in ln
            continue
        if not active:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
        if m:
            table.append((int(m.group(1), 16), cur, m.group(2).strip()))
    return table


def main():
    rep, so, ksub = sys.argv[1:4]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[hdr_i]
    ci = hdr.index("Instructions Executed")
    ti = hdr.index("Thread Instructions Executed")
    si = hdr.index("# Samples")
    data = [r for r in rows[hdr_i + 1:] if len(r) > ti]
    table = line_table(so, ksub)
    assert len(table) == len(data), (len(table), len(data))
    by_line = defaultdict(lambda: [0, 0, 0])
    by_file = defaultdict(lambda: [0, 0, 0])
    tot = [0, 0, 0]
    for (off, loc, text), r in zip(table, data):
        w, t, s = int(r[ci]), int(r[ti]), int(r[si])
        for d in (by_line[loc], by_file[loc[0] if loc else None], tot):
            d[0] += w
            d[1] += t
            d[2] += s
    print("total warp-inst %d thread-inst %d avg active %.2f samples %d  SASS instrs %d" %
          (tot[0], tot[1], tot[1] / max(1, tot[0]), tot[2], len(table)))
    print("\nper file:")
    for f, (w, t, s) in sorted(by_file.items(), key=lambda kv: -kv[1][0]):
        print("  %-22s warp-inst %5.1f%%  avg-active %5.2f  samples %5.1f%%" %
              (f, 100.0 * w / tot[0], t / max(1, w), 100.0 * s / max(1, tot[2])))
    print("\ntop lines:")
    for loc, (w, t, s) in sorted(by_line.items(), key=lambda kv: -kv[1][0])[:top]:
        print("  %-22s:%-4d warp-inst %5.2f%%  avg-active %5.2f  samples %5.2f%%" %
              (loc[0], loc[1], 100.0 * w / tot[0], t / max(1, w), 100.0 * s / max(1, tot[2])))


if __name__ == "__main__":
    main()
