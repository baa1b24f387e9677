#!/usr/bin/env python
"""Device-resident simulation rate (reactions/s, CUDA events, 3 timed launches after 1 warm-up) for the
single-locus configurations of BASELINE.json: cfg1, cfg2, cfg3.  One JSON line per configuration."""
import json
import os
import sys

import numpy as np

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import math  # noqa: E402

from sonics_b200 import driver, engine as eng  # noqa: E402

CFG1 = "8|5;9|113;10|89"
CFG2 = "5|1;6|1;7|5;8|11;9|20;10|24;11|2;12|1"
CFG3 = ";".join(
    "%d|%d" % (a, int(5 + 400 * math.exp(-0.5 * ((a - 28) / 6) ** 2) + 300 * math.exp(-0.5 * ((a - 36) / 5) ** 2)))
    for a in range(10, 50))


def main():
    e = eng.Engine(0)
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 2000000
    cases = [("cfg1", CFG1, {}), ("cfg2", CFG2, {}), ("cfg3", CFG3, dict(n_cycles=30, capture_cycle=16))]
    for name, spec, kw in cases:
        # the CLI's own pre-filter: alleles + automatic noise coefficient (sonics:407-416)
        kind, (reads, nc) = driver.prefilter(("sample", "Block", ".", spec),
                                             {"one_allele_threshold": 45, "noise_threshold": 0.05},
                                             {"monte_carlo": True})
        nc = float(np.float32(nc))
        params = eng.make_params(**kw)
        table = e.build_table_from_arrays(params, [reads], [nc])
        out = e.simulate(params, table, n, seed=1)
        e.sync()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        for i in range(3):
            e.simulate(params, table, n, seed=2 + i, out=out)
        t1.record()
        torch.cuda.synchronize()
        ms = t0.elapsed_time(t1) / 3
        lnl = out["lnL"][0].cpu().numpy()
        print(json.dumps({"config": name, "reactions_per_s": n / ms * 1e3, "ms_per_launch": ms, "reactions": n,
                          "window_bins": int(table.max_window), "noise_coef": nc,
                          "sentinel_fraction": float((lnl == -999999.0).mean())}))


if __name__ == "__main__":
    main()
