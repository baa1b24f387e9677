#!/usr/bin/env python
"""One K1 launch of a single-locus configuration (for ncu): python tools/prof_k1.py [cfg2] [reactions] [kernel]."""
import math
import os
import sys

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)

from sonics_b200 import engine as eng  # noqa: E402
from sonics_b200.sonics import get_alleles  # noqa: E402

CFG = {"cfg1": ("8|5;9|113;10|89", {}),
       "cfg2": ("5|1;6|1;7|5;8|11;9|20;10|24;11|2;12|1", {}),
       "cfg3": (";".join("%d|%d" % (a, int(5 + 400 * math.exp(-0.5 * ((a - 28) / 6) ** 2) +
                                          300 * math.exp(-0.5 * ((a - 36) / 5) ** 2))) for a in range(10, 50)),
                dict(n_cycles=30, capture_cycle=16))}
name = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1000000
kernel = sys.argv[3] if len(sys.argv) > 3 else "auto"
spec, kw = CFG[name]
e = eng.Engine(0)
e.set_kernel(kernel)
params = eng.make_params(**kw)
reads = get_alleles(spec)[0]
nz = reads[reads > 0]
frac = nz.min() / reads.sum() * len(nz)
nc = float(nz.min() / reads.sum()) if frac / (frac + 1) < 0.05 else 0.0  # sonics:407-416
table = e.build_table_from_arrays(params, [reads], [nc])
out = e.simulate(params, table, n, seed=1)
e.sync()
print(name, n, kernel, float((out["lnL"] == -999999.0).double().mean()))
