#!/usr/bin/env python
"""Print K1's control-flow counters for the cfg2 locus (needs a GPU)."""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch  # noqa: E402

from sonics_b200 import engine as eng  # noqa: E402
from sonics_b200 import sonics as drop_in  # noqa: E402
from sonics_b200._abi import check  # noqa: E402

gt = sys.argv[1] if len(sys.argv) > 1 else "5|1;6|1;7|5;8|11;9|20;10|24;11|2;12|1"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 200000
e = eng.Engine(0)
p = eng.make_params()
t = e.build_table_from_arrays(p, [drop_in.get_alleles(gt)[0]], [0.0])
cnt = torch.zeros(16, dtype=torch.int64, device=e.device)
check(e.lib.sonics_debug_counters(e.ctx, cnt.data_ptr()))
e.simulate(p, t, n, seed=1)
e.sync()
check(e.lib.sonics_debug_counters(e.ctx, None))
c = cnt.cpu().numpy().astype(float)
warps = c[12]
names = ["ADV", "BTRS", "INV0", "BOUND", "FULL", "POIS", "EXIT"]
print("warps %d, passes/warp %.1f, passes per simulation-lane %.1f" % (warps, c[0] / warps, c[0] * 32 / n))
for i, nm in enumerate(names):
    print("  lanes in %-5s per pass: %.2f" % (nm, c[1 + i] / c[0]))
print("ADVANCE: lanes entering per pass %.2f, inner trips per pass %.2f" % (c[9] / c[0], c[8] / c[0]))
print("FULL: runs per pass %.3f, lanes per run %.2f" % (c[10] / c[0], c[11] / max(1, c[10])))
