#!/usr/bin/env python
"""Schedule counters of the lockstep kernel (diagnostics build, -DSONICS_LS_STATS).

  python tools/ls_stats.py build          compile sonics_b200/variants/ls_stats.so (here, no GPU needed)
  python tools/ls_stats.py run [cfg2] [reactions]   (on the GPU box) one launch, counters as JSON
"""
import ctypes
import json
import os
import subprocess
import sys

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
VAR = os.path.join(ROOT, "sonics_b200", "variants", "ls_stats.so")


def build():
    from sonics_b200 import build as b
    os.makedirs(os.path.dirname(VAR), exist_ok=True)
    cmd = [os.environ.get("NVCC", "nvcc")] + b.NVCC_FLAGS + ["-DSONICS_LS_STATS", "-o", VAR] + b.SOURCES
    subprocess.check_call(cmd, cwd=b.CSRC)
    print(VAR)


def run(name, n):
    os.environ["SONICS_B200_LIB"] = VAR
    import numpy as np
    from sonics_b200 import engine as eng
    from sonics_b200.sonics import get_alleles
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import k1_ab
    e = eng.Engine(0)
    e.set_kernel("lockstep")
    if name.startswith("vcf"):
        # vcf:<loci>:<pair|slot>  -- a cfg4-style table, n simulations per genotype
        import synth_vcf
        from sonics_b200 import driver
        _, n_loci, mode = name.split(":")
        loci = synth_vcf.make_genotypes(int(n_loci), 8, seed=20261019, cov=(50, 1000))
        pre = {"one_allele_threshold": 45, "noise_threshold": 0.05}
        work = [p for k, p in (driver.prefilter(t, pre, {"monte_carlo": False})
                               for t in synth_vcf.genotype_strings(loci)) if k == "sim"]
        params = eng.make_params()
        table = e.build_table_from_arrays(params, [w[0] for w in work], [float(np.float32(w[1])) for w in work])
        e.lib.sonics_set_sort_mode(e.ctx, 1 if mode == "slot" else 0)
    else:
        spec, kw = {"cfg1": (k1_ab.CFG1, {}), "cfg2": (k1_ab.CFG2, {}),
                    "cfg3": (k1_ab.CFG3, dict(n_cycles=30, capture_cycle=16))}[name]
        params = eng.make_params(**kw)
        reads = get_alleles(spec)[0]
        nz = reads[reads > 0]
        frac = nz.min() / reads.sum() * len(nz)
        nc = float(nz.min() / reads.sum()) if frac / (frac + 1) < 0.05 else 0.0
        table = e.build_table_from_arrays(params, [reads], [nc])
    buf = (ctypes.c_uint64 * 128)()
    e.lib.sonics_debug_ls_stats.argtypes = [ctypes.c_void_p]
    e.simulate(params, table, n, seed=1)
    e.sync()
    e.lib.sonics_debug_ls_stats(buf)  # clear
    import time
    t0 = time.perf_counter()
    e.simulate(params, table, n, seed=2)
    e.sync()
    ms = (time.perf_counter() - t0) * 1e3
    e.lib.sonics_debug_ls_stats(buf)
    n = n * table.n_gt
    c = np.array(list(buf), dtype=np.float64)
    nb = c[0]
    d = {
        "workload": name, "reactions": n, "batches": nb, "ms_instrumented": ms,
        "sweep_steps_per_batch": c[13] / nb, "lanes_in_per_sweep_step": c[14] / max(c[13], 1),
        "draw_slots_per_batch": c[1] / nb,
        "slots_by_stage_per_batch": [c[48 + i] / nb for i in range(3)],
        "lanes_drawing_by_stage": [c[52 + i] / max(c[48 + i], 1) for i in range(3)],
        "btrs_slots_per_batch": c[2] / nb, "inv_slots_per_batch": c[3] / nb, "mixed_slots_per_batch": c[4] / nb,
        "btrs_lanes_per_btrs_slot": c[5] / max(c[2], 1), "inv_lanes_per_inv_slot": c[6] / max(c[3], 1),
        "btrs_rounds_per_btrs_slot": c[7] / max(c[2], 1), "lanes_per_btrs_round": c[8] / max(c[7], 1),
        "rounds_with_undecided_frac": c[9] / max(c[7], 1), "undecided_per_trial": c[10] / max(c[8], 1),
        "inv_max_steps_per_inv_slot": c[11] / max(c[3], 1), "inv_mean_steps_per_inv_lane": c[12] / max(c[6], 1),
        "rounds_hist": [c[16 + i] / max(c[2], 1) for i in range(16)],
        "btrs_lanes_by_log2_npq": {str(k): c[32 + k] / max(c[5], 1) for k in range(3, 16)},
    }
    print(json.dumps(d))


if __name__ == "__main__":
    if sys.argv[1] == "build":
        build()
    else:
        run(sys.argv[2] if len(sys.argv) > 2 else "cfg2", int(sys.argv[3]) if len(sys.argv) > 3 else 200000)
