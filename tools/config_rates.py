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

from oracle import oracle as orc  # noqa: E402  (parse_reads / auto_noise_coef only: input preparation)
from sonics_b200 import engine as eng  # noqa: E402
from tests.util import CFG1, CFG2, CFG3  # noqa: E402


def main():
    e = eng.Engine(0)
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 2000000
    cases = [("cfg1", CFG1, 0.0, {}), ("cfg2", CFG2, 0.0, {}),
             ("cfg3", CFG3, None, dict(n_cycles=30, capture_cycle=16))]
    for name, spec, nc, kw in cases:
        reads = orc.parse_reads(spec)
        if nc is None:
            nc = float(np.float32(orc.auto_noise_coef(reads)))
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
