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
e = eng.Engine(0)
loci = synth_vcf.make_genotypes(n_loci, 8, seed=20261019, cov=(50, 1000))
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
print("genotypes %d in %.3f s = %.0f /s; reactions %d" % (len(work), dt, len(work) / dt, int(v["run_reps"].sum())))
