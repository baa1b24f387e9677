#!/usr/bin/env python
"""Run the cfg4 genotyping workload once (needs a GPU); used under ncu for the launch list."""
import os
import sys
import time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import numpy as np  # noqa: E402
import synth_vcf  # noqa: E402

from sonics_b200 import driver, engine as eng  # noqa: E402

n_loci = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
n_samples = int(sys.argv[2]) if len(sys.argv) > 2 else 8
noisy = float(sys.argv[3]) if len(sys.argv) > 3 else 0.0
cov = (2000, 10000) if noisy > 0 else (50, 1000)
e = eng.Engine(0)
if float(os.environ.get("SONICS_FAST_SAMPLERS", "0") or 0) > 0:   # the opt-in approximate sampler
    e.set_samplers("fast", float(os.environ["SONICS_FAST_SAMPLERS"]))
loci = synth_vcf.make_genotypes(n_loci, n_samples, seed=20261019, cov=cov, noisy=noisy)
constants = {"one_allele_threshold": 45, "noise_threshold": 0.05}
work = [p for k, p in (driver.prefilter(t, constants, {"monte_carlo": False}) for t in synth_vcf.genotype_strings(loci))
        if k == "sim"]
params = eng.make_params()
off, al, rd = eng.to_csr([w[0] for w in work])
noise = np.array([w[1] for w in work], dtype=np.float32)
e.genotype_batch(params, off[:101], al[:off[100]], rd[:off[100]], noise[:100], 1000, seed=1)
t0 = time.perf_counter()
v = e.genotype_batch(params, off, al, rd, noise, 1000, seed=20261019)
dt = time.perf_counter() - t0
from sonics_b200.sonics import filter_string
f, c = np.unique(v["filter"], return_counts=True)
print("genotypes %d in %.3f s = %.0f /s; reactions %d (%.3e /s); repeats %s; filters %s" % (
    len(work), dt, len(work) / dt, int(v["run_reps"].sum()), v["run_reps"].sum() / dt,
    {int(k): int((v["run_reps"] == k).sum()) for k in (100, 500, 1000)},
    {filter_string(int(a)) or "(none)": int(b) for a, b in zip(f, c)}))
