#!/usr/bin/env python
"""Static audit of a kernel's SASS for constructs that cost many instructions per source operation.

usage: sass_audit.py <libsonics_b200.so> <kernel-substring> [report.ncu-rep]

Per source line (nvdisasm -g line info): slow-path CALLs (IEEE division, sqrt, 64-bit division), range-scaling
sequences around MUFU ops (the |x| < 1.175e-38 test that __fdividef / __logf / sqrtf expand to), register
spills (LDL / STL) and 64-bit integer adds.  With an ncu report (--set full --import-source on) of the same
build the executed warp-instruction share of each line is added, so cold code can be ignored.  This is how
the v4.7-v4.9 changes of the PCR kernel were found."""
import csv
import io
import os
import re
import subprocess
import sys
from collections import defaultdict

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import ncu_lines  # noqa: E402

CLASSES = [
    ("call", re.compile(r"\bCALL")),
    ("range-scaling", re.compile(r"1\.175494350822287508e-38")),
    ("mufu", re.compile(r"\bMUFU")),
    ("spill", re.compile(r"\b(LDL|STL)")),
    ("int64", re.compile(r"IADD3\.X|IMAD\.X|IMAD\.WIDE|LEA\.HI\.X")),
    ("f64", re.compile(r"\b(DADD|DMUL|DFMA|DSETP|MUFU\.RCP64H)")),
]


def main():
    so, ksub = sys.argv[1:3]
    rep = sys.argv[3] if len(sys.argv) > 3 else None
    table = ncu_lines.line_table(so, ksub)
    executed = None
    if rep:
        out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(out)))
        hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
        ci = rows[hi].index("Instructions Executed")
        data = [r for r in rows[hi + 1:] if len(r) > ci]
        if len(data) == len(table):
            executed = [int(r[ci]) for r in data]
        else:
            print("note: report has %d instructions, the library %d -- different build, counts ignored" %
                  (len(data), len(table)))
    per_line = defaultdict(lambda: defaultdict(int))
    weight = defaultdict(int)
    total = sum(executed) if executed else 0
    for i, (off, loc, text) in enumerate(table):
        if executed:
            weight[loc] += executed[i]
        for name, pat in CLASSES:
            if pat.search(text):
                per_line[loc][name] += 1
    print("%d SASS instructions in kernels matching %r" % (len(table), ksub))
    print("%-28s %8s  %s" % ("source line", "exec %" if executed else "", "constructs"))
    keyf = (lambda kv: -weight[kv[0]]) if executed else (lambda kv: str(kv[0]))
    for loc, c in sorted(per_line.items(), key=keyf):
        if set(c) == {"mufu"}:
            continue  # a bare MUFU is the cheap form
        share = "%7.2f%%" % (100.0 * weight[loc] / total) if executed else ""
        name = "%s:%d" % loc if loc else "?"
        print("%-28s %8s  %s" % (name, share, ", ".join("%s x%d" % kv for kv in sorted(c.items()))))


if __name__ == "__main__":
    main()
