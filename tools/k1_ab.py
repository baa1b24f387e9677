#!/usr/bin/env python
"""A/B of the two forms of K1 (persistent lanes vs lockstep on sorted units), device-resident, CUDA events.

  python tools/k1_ab.py [reactions]           cfg1 / cfg2 / cfg3 single-locus launches and one genotyping-shaped
                                              launch (2000 genotypes x 100 / x 400 simulations)
  SONICS_AB_BITS="auto;0,0,0;6,4,2"           sort-key bit splits to try for the lockstep form
One JSON line per (workload, kernel, bits)."""
import json
import math
import os
import sys

import numpy as np

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

import torch  # noqa: E402

import synth_vcf  # noqa: E402
from sonics_b200 import driver, engine as eng  # noqa: E402
from sonics_b200._abi import check  # noqa: E402

CFG1 = "8|5;9|113;10|89"
CFG2 = "5|1;6|1;7|5;8|11;9|20;10|24;11|2;12|1"
CFG3 = ";".join(
    "%d|%d" % (a, int(5 + 400 * math.exp(-0.5 * ((a - 28) / 6) ** 2) + 300 * math.exp(-0.5 * ((a - 36) / 5) ** 2)))
    for a in range(10, 50))


def timed(e, params, table, n, reps=3, **kw):
    out = e.simulate(params, table, n, seed=1, **kw)
    e.sync()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for i in range(reps):
        e.simulate(params, table, n, seed=2 + i, out=out, **kw)
    t1.record()
    torch.cuda.synchronize()
    return t0.elapsed_time(t1) / reps, out


def main():
    e = eng.Engine(0)
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 4000000
    only = set(sys.argv[2].split(",")) if len(sys.argv) > 2 else None
    bits_list = os.environ.get("SONICS_AB_BITS", "auto").split(";")
    pre = {"one_allele_threshold": 45, "noise_threshold": 0.05}
    cases = []
    for name, spec, kw, nn in (("cfg1", CFG1, {}, n), ("cfg2", CFG2, {}, n),
                               ("cfg3", CFG3, dict(n_cycles=30, capture_cycle=16), max(n // 8, 1000))):
        kind, (reads, nc) = driver.prefilter(("sample", "Block", ".", spec), pre, {"monte_carlo": True})
        cases.append((name, [reads], [float(np.float32(nc))], kw, nn))
    loci = synth_vcf.make_genotypes(int(os.environ.get("SONICS_AB_LOCI", "250")), 8, seed=20261019, cov=(50, 1000))
    work = [p for k, p in (driver.prefilter(t, pre, {"monte_carlo": False}) for t in synth_vcf.genotype_strings(loci))
            if k == "sim"]
    for reps in (100, 400):
        cases.append(("vcf_%d" % reps, [w[0] for w in work], [float(np.float32(w[1])) for w in work], {}, reps))
    for name, reads, noise, kw, nn in cases:
        if only and name not in only:
            continue
        params = eng.make_params(**kw)
        table = e.build_table_from_arrays(params, reads, noise)
        total = nn * table.n_gt
        ref = None
        for kernel in ("persistent", "lockstep"):
            for bits in (bits_list if kernel == "lockstep" else ["-"]):
                e.set_kernel(kernel)
                if kernel == "lockstep":
                    b = (-1, -1, -1) if bits == "auto" else tuple(int(x) for x in bits.split(","))
                    check(e.lib.sonics_set_sort_bits(e.ctx, *b))
                for mode in ((0, 1) if kernel == "lockstep" and table.n_gt > 1 else (0,)):
                    check(e.lib.sonics_set_sort_mode(e.ctx, mode))
                    ms, out = timed(e, params, table, nn)
                    if mode == 1:
                        print(json.dumps({"workload": name, "kernel": kernel, "sort_bits": bits, "sort_mode": "slot",
                                          "reactions": total, "ms": ms, "reactions_per_s": total / ms * 1e3}), flush=True)
                check(e.lib.sonics_set_sort_mode(e.ctx, 0))
                ms, out = timed(e, params, table, nn)
                h = out["lnL"].cpu().numpy().tobytes()
                same = ref is None or h == ref
                ref = ref or h
                print(json.dumps({"workload": name, "kernel": kernel, "sort_bits": bits, "reactions": total,
                                  "ms": ms, "reactions_per_s": total / ms * 1e3, "same_as_first": bool(same)}), flush=True)
        check(e.lib.sonics_set_sort_bits(e.ctx, -1, -1, -1))
        e.set_kernel("auto")


if __name__ == "__main__":
    main()
