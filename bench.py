#!/usr/bin/env python
"""Benchmark of the SONiCS PCR-stutter Monte-Carlo hot path on B200.

    python bench.py --gpus N --steps K --warmup W [--impl reference]

Workload (BASELINE.json configs[1], the configuration the metric is quoted on): the single
locus "5|1;6|1;7|5;8|11;9|20;10|24;11|2;12|1", default CLI constants, fixed-count simulation,
1e6 simulations per genotype group x 8 groups = 8e6 PCR reactions per step per GPU.
A step = one pass of the hot path (K1: priors -> PCR cycles -> capture -> lnL) over that batch.
Under torchrun (N > 1) every rank simulates its own disjoint range of simulation indices of the
same locus (weak scaling, no collective on the data path); rank 0 prints ONE JSON line.

  value  simulated PCR reactions/s, whole job, inputs resident in HBM, CUDA-event timed
  e2e    same metric through the public host API as a stream of jobs (genotype string in host memory -> table
         build + H2D -> simulate -> lnL/group arrays copied back to pinned host memory; the copy of step i overlaps
         the simulation of step i + 1, all copies inside the timed region)
  roofline        the bounding roofline: warp instructions issued per second against
                  148 SMs x 4 schedulers x SM clock.  Instructions per reaction come from the committed ncu
                  capture (profiles/k1_counters.json, guarded by a hash of the kernel sources: a stale file
                  gives frac = null), the rate is measured live.  lanes_active_of_32 and achieved_algorithmic
                  (thread-instructions a simulation needs / thread-instruction peak) say how much of the
                  issued work is useful; traffic = DRAM bytes of every kernel of the step
  hbm_roofline    the HBM view (10 B/reaction of mandatory traffic): tiny by construction
  configs         device-resident rates of cfg1 / cfg3, cfg2 with the per-simulation readout (identity / r^2)
                  switched on -- the work the reference's one_repeat() does -- and cfg2 with the opt-in approximate
                  sampler (SONICS_SAMPLERS_FAST, min_npq 100): beside `value`, which is always the exact mode
  genotyping      the other half of BASELINE.json's metric, loci genotyped/s: cfg4 through the C-ABI with
                  host buffers, a cfg5-style noisy VCF file to file across all ranks (strong scaling: seconds,
                  genotypes/s, sha256 of the output) and, at N = 1, the reference's monte_carlo() on the host
                  cores over a bounded subset of the same genotypes
  cpu_baseline    the reference's own one_repeat() (compiled from /root/reference/sonics.pyx into
                  oracle/_ref by oracle/build_ref.py) on the box's host cores, bounded sample

--impl reference times that CPU implementation instead (all host cores, bounded sample per step).
"""
import argparse
import json
import multiprocessing as mp
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CFG2 = "5|1;6|1;7|5;8|11;9|20;10|24;11|2;12|1"
SIMS_PER_STEP = 8_000_000
BYTES_PER_REACTION = 10  # fp64 lnL + u16 group id written per reaction
METRIC = "simulated_pcr_reactions_per_s"
UNIT = "reactions/s"


# --------------------------------------------------------------------------- CPU arm
def _ref_worker(args):
    n, seed = args
    sys.path.insert(0, os.path.join(ROOT, "oracle", "_ref"))
    import numpy as np
    import sonics as ref  # the compiled reference

    alleles, max_allele, _ = ref.get_alleles(CFG2)
    constants = dict(random=False, n_cycles=24, capture_cycle=13, up_preference=False, floor=5, lnL_threshold=2.3,
                     start_copies=300000, genotype=CFG2, padjust=0.001, noise_threshold=0.05,
                     one_allele_threshold=45, max_allele=max_allele, genotype_total=sum(alleles), alleles=alleles,
                     noise_coef=0)
    ranges = ((0, 0.1), (0, 0.1), (-0.25, 0.25), (0.001, 0.1))
    np.random.seed(seed)
    t0 = time.perf_counter()
    for _ in range(n):
        ref.one_repeat(constants, ranges, 100)
    return time.perf_counter() - t0


def _port_worker(args):
    n, seed = args
    from oracle import oracle as orc
    cfg = orc.make_cfg(rng_mode=0, with_readout=True)
    t0 = time.perf_counter()
    orc.run(cfg, orc.parse_reads(CFG2), seed, n)
    return time.perf_counter() - t0


def have_ref():
    d = os.path.join(ROOT, "oracle", "_ref")
    return os.path.isdir(d) and any(f.startswith("sonics.") and f.endswith(".so") for f in os.listdir(d))


def cpu_sample(kind, per_worker, cores, seed):
    """Run per_worker one_repeat() calls on each of `cores` processes; returns (reactions, seconds)."""
    worker = _ref_worker if kind == "reference" else _port_worker
    ctx = mp.get_context("fork")
    with ctx.Pool(cores) as pool:
        times = pool.map(worker, [(per_worker, seed + i) for i in range(cores)])
    # the workers run concurrently; the slowest one bounds the sample (process start-up and imports excluded)
    return per_worker * cores, max(times)


def cpu_baseline(target_s=12.0):
    kind = "reference" if have_ref() else "port"
    cores = os.cpu_count() or 1
    # calibrate on a short run, then size the sample for ~target_s seconds
    n0 = 40 if kind == "reference" else 200
    r, s = cpu_sample(kind, n0, cores, 1)
    per_worker = max(n0, int(n0 * target_s / max(s, 1e-3)))
    r, s = cpu_sample(kind, per_worker, cores, 100)
    single = cpu_sample(kind, max(20, per_worker // 8), 1, 200)
    return {"value": r / s, "unit": UNIT, "cores": cores, "kind": kind,
            "sample": "%d one_repeat() calls of the cfg2 locus on %d processes (%.1f s)" % (r, cores, s),
            "single_core_value": single[0] / single[1]}


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    kind = "reference" if have_ref() else "port"
    cores = os.cpu_count() or 1
    per_worker = 250 if kind == "reference" else 1200
    for i in range(args.warmup):
        cpu_sample(kind, max(4, per_worker // 10), cores, 7 + i)
    total_r, total_s = 0, 0.0
    for i in range(args.steps):
        r, s = cpu_sample(kind, per_worker, cores, 1000 + i)
        total_r += r
        total_s += s
    v = total_r / total_s
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total_s / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "cfg2 single locus %s, default constants; each step a bounded sample of %d "
                               "one_repeat() calls (the GPU arm runs %d per step per GPU)" %
                               (CFG2, per_worker * cores, SIMS_PER_STEP)},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": kind,
                         "sample": "%d one_repeat() calls per step on %d processes" % (per_worker * cores, cores)},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


# --------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for row in self.rows:
            f = [x.strip() for x in row.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for name, val in zip(names, f[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# --------------------------------------------------------------------------- GPU arm
def load_inst_per_reaction():
    p = os.path.join(ROOT, "profiles", "k1_counters.json")
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f)
    return None


def genotyping_throughput(e, eng, drop_in, n_loci=10000, n_samples=8):
    """Second half of BASELINE.json's metric: loci genotyped/s (config 4: synthetic lobSTR-style VCF,
    10k loci x 8 samples, default filters, the full repeat-until-pass loop).  Measured end to end
    through sonics_genotype_batch with HOST buffers (CSR in, verdicts out), after one warm-up call."""
    import numpy as np
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import synth_vcf
    from sonics_b200 import driver
    loci = synth_vcf.make_genotypes(n_loci, n_samples, seed=20261019, cov=(50, 1000))
    constants = {"one_allele_threshold": 45, "noise_threshold": 0.05}
    work, items = [], []
    for t in synth_vcf.genotype_strings(loci):
        kind, payload = driver.prefilter(t, constants, {"monte_carlo": False})
        if kind == "sim":
            work.append(payload)
            if len(items) < 4096:
                items.append((t[3], payload[0], float(payload[1])))
    params = eng.make_params()
    off, al, rd = eng.to_csr([w[0] for w in work])
    noise = np.array([w[1] for w in work], dtype=np.float32)
    e.genotype_batch(params, off[:1001], al[:off[1000]], rd[:off[1000]], noise[:1000], 1000, seed=1)  # warm-up
    l0 = e.launches
    t0 = time.perf_counter()
    v = e.genotype_batch(params, off, al, rd, noise, 1000, seed=20261019)
    dt = time.perf_counter() - t0
    n_launch = int(e.launches - l0)
    reps = v["run_reps"].astype(np.int64)
    filt, cnt = np.unique(v["filter"], return_counts=True)
    # the same loop with the opt-in approximate sampler (beside the exact figure, never instead of it)
    e.set_samplers("fast", 100.0)
    try:
        t0 = time.perf_counter()
        vf = e.genotype_batch(params, off, al, rd, noise, 1000, seed=20261019)
        dt_fast = time.perf_counter() - t0
    finally:
        e.set_samplers("exact")
    same_call = float(np.mean((vf["allele_lo"] == v["allele_lo"]) & (vf["allele_hi"] == v["allele_hi"])))
    # the same job file to file: VCF text -> native ingest + pre-filters -> GPU loop -> sonics_out.txt
    import tempfile
    with tempfile.TemporaryDirectory() as d:
        vcf_path = os.path.join(d, "cfg4.vcf")
        synth_vcf.write_vcf(vcf_path, loci, n_samples)
        args = driver.build_parser().parse_args(["-m", vcf_path, "-o", d, "--seed", "20261019"])
        c, r, o = driver.dicts_from_args(args)
        import logging
        logging.disable(logging.WARNING)
        t1 = time.perf_counter()
        driver.run_sonics(vcf_path, c, r, o, seed=20261019)
        file_dt = time.perf_counter() - t1
        logging.disable(logging.NOTSET)
        n_lines = sum(1 for _ in open(os.path.join(d, "sonics_out.txt"))) - 1
    return {"value": len(work) / dt, "unit": "genotypes/s", "genotypes": len(work), "seconds": dt,
            "reactions": int(reps.sum()), "reactions_per_s": float(reps.sum() / dt),
            "repeats_hist": {str(k): int((reps == k).sum()) for k in (100, 500, 1000)},
            "filter_hist": {drop_in.filter_string(int(f)) or "(none)": int(c) for f, c in zip(filt, cnt)},
            "gpu_launches": n_launch,
            "fast_samplers": {"value": len(work) / dt_fast, "unit": "genotypes/s", "seconds": dt_fast, "min_npq": 100,
                              "same_genotype_as_exact": same_call,
                              "what": "the same call with SONICS_SAMPLERS_FAST (opt-in; `value` is the exact mode)"},
            "file_to_file": {"value": n_lines / file_dt, "unit": "genotypes/s", "seconds": file_dt, "rows": n_lines,
                             "what": "sonics -m cfg4.vcf: native VCF ingest + pre-filters, GPU loop, native row writer"},
            "workload": "cfg4: synthetic VCF %d loci x %d samples, coverage 50-1000, -r 1000, default filters; "
                        "host CSR in -> verdicts out through sonics_genotype_batch" % (n_loci, n_samples)}, items


K1_SOURCES = ["sim_lockstep.cuh", "sim_kernel.cuh", "samplers.cuh", "rng.cuh", "finish.cuh", "lnl.cuh"]


def kernel_sha():
    """Hash of the K1 sources (comments and white space stripped): profiles/k1_counters.json is only valid for the
    code it was captured from."""
    import hashlib
    import re
    h = hashlib.sha256()
    for f in K1_SOURCES:
        with open(os.path.join(ROOT, "sonics_b200", "csrc", f), "r") as fh:
            src = fh.read()
        src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
        src = re.sub(r"//[^\n]*", "", src)
        h.update(re.sub(r"\s+", "", src).encode())
    return h.hexdigest()[:16]


def _geno_ref_worker(args):
    """monte_carlo() of the compiled reference (or of the oracle port) over a list of pre-filtered genotypes."""
    items, seed, kind = args
    import numpy as np
    if kind == "reference":
        sys.path.insert(0, os.path.join(ROOT, "oracle", "_ref"))
        import sonics as ref
    else:
        from oracle import oracle as orc
        from oracle import stats
    ranges = ((0, 0.1), (0, 0.1), (-0.25, 0.25), (0.001, 0.1))
    options = dict(repetitions=1000, out_path=".", strict=1, file_name="x", vcf_mode=True, name="s", block="b",
                   processes=0, verbose=False, save_report=False, monte_carlo=False, add_ratios="", min_sim=25)
    np.random.seed(seed)
    t0 = time.perf_counter()
    reps = 0
    for gt, reads, nc in items:
        if kind == "reference":
            constants = dict(random=False, n_cycles=24, capture_cycle=13, up_preference=False, floor=5,
                             lnL_threshold=2.3, start_copies=300000, genotype=gt, padjust=0.001, noise_threshold=0.05,
                             one_allele_threshold=45, max_allele=500, genotype_total=int(reads.sum()), alleles=reads,
                             noise_coef=nc)
            row = ref.monte_carlo(1000, constants, ranges, options)
        else:
            cfg = orc.make_cfg(noise_coef=nc, rng_mode=0, with_readout=True)
            state = {"k": 0}

            def draw(n, cfg=cfg, reads=reads, state=state):
                state["k"] += 1
                recs = orc.run(cfg, reads, seed * 1000 + state["k"], n)
                return [dict(identity=r["identity"], r_squared=r["r_squared"], lnL=r["lnL"], label=orc.label(r))
                        for r in recs]
            row, _ = stats.monte_carlo(1000, draw)
        reps += int(row.split("\t")[8])
    return time.perf_counter() - t0, reps


def genotyping_cpu_baseline(work_items, target_s=15.0):
    """The reference's monte_carlo() (sonics.pyx:47-269) over the first genotypes of the cfg4 VCF, one fork-pool
    worker per host core (the reference's own -p strategy, sonics:216-221), sized for ~target_s seconds."""
    kind = "reference" if have_ref() else "port"
    cores = os.cpu_count() or 1
    ctx = mp.get_context("fork")
    # calibrate on one genotype per core, then take as many as fit the time budget (at least two per core)
    with ctx.Pool(cores) as pool:
        t = pool.map(_geno_ref_worker, [([work_items[i]], 1 + i, kind) for i in range(cores)])
    per = max(1e-3, sum(x[0] for x in t) / cores)
    per_core = int(max(2, min(40, target_s / per)))
    n = min(len(work_items), cores * per_core)
    jobs = [(work_items[i::cores][:per_core], 100 + i, kind) for i in range(cores)]
    with ctx.Pool(cores) as pool:
        t = pool.map(_geno_ref_worker, jobs)
    secs = max(x[0] for x in t)
    n = sum(len(j[0]) for j in jobs)
    return {"value": n / secs, "unit": "genotypes/s", "cores": cores, "kind": kind,
            "sample": "monte_carlo() on the first %d genotypes of the cfg4 VCF, %d per process on %d processes "
                      "(%.1f s); a full-file figure would be this rate extrapolated" % (n, per_core, cores, secs),
            "reactions": int(sum(x[1] for x in t)), "seconds": secs}


def cfg5_subset_vcf(path, distinct_loci=2500, copies=40, n_samples=32):
    """cfg5 input (BASELINE.json configs[4]: 100k loci x 32 samples, coverage 2000-10000, 10 % noisy profiles ->
    no_success / MWU_test at 1000 repetitions) as `copies` x `distinct_loci` loci x 32 samples: the distinct
    profiles are generated once and written `copies` times under different locus names (different RNG keys),
    which keeps the generator out of the benchmark's run time."""
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import synth_vcf
    loci = synth_vcf.make_genotypes(distinct_loci, n_samples, seed=20261019, cov=(2000, 10000), noisy=0.1)
    columns = synth_vcf.sample_columns(loci)
    for c in range(copies):
        synth_vcf.write_vcf(path, loci, n_samples, start=c * distinct_loci, header=(c == 0), columns=columns,
                            mode="w" if c == 0 else "a")
    return distinct_loci * copies * n_samples


def genotyping_file_to_file(rank, world, barrier, workdir, copies=40):
    """Strong scaling of the genotyping job: one cfg5-style VCF -> sonics_out.txt across all ranks
    (driver.run_sonics under torchrun: round-robin byte pieces, no collective on the data path)."""
    import hashlib
    import logging
    from sonics_b200 import driver
    vcf = os.path.join(workdir, "cfg5_subset.vcf")
    warm = os.path.join(workdir, "warm.vcf")
    n_gt = 0
    if rank == 0:
        n_gt = cfg5_subset_vcf(vcf, copies=copies)
        cfg5_subset_vcf(warm, distinct_loci=64, copies=1)
    barrier()
    logging.disable(logging.WARNING)
    out = {}
    for name, path in (("warm", warm), ("run", vcf)):
        d = os.path.join(workdir, "out_" + name)
        args = driver.build_parser().parse_args(["-m", path, "-o", d, "--seed", "20261019"])
        c, r, o = driver.dicts_from_args(args)
        barrier()
        t0 = time.perf_counter()
        driver.run_sonics(path, c, r, o, seed=20261019)
        barrier()
        out[name] = (time.perf_counter() - t0, os.path.join(d, "sonics_out.txt"))
    logging.disable(logging.NOTSET)
    secs, path = out["run"]
    if rank != 0:
        return None
    h = hashlib.sha256()
    rows = -1
    with open(path, "rb") as f:
        for blk in iter(lambda: f.read(1 << 22), b""):
            h.update(blk)
            rows += blk.count(b"\n")
    return {"value": rows / secs, "unit": "genotypes/s", "n_gpus": world, "seconds": secs, "rows": rows,
            "output_sha256": h.hexdigest(), "scaling": "strong",
            "workload": "cfg5 VCF: %d loci (2500 distinct profiles x %d) x 32 samples, coverage "
                        "2000-10000, 10 %% noisy profiles, -r 1000, default filters; VCF file -> sonics_out.txt "
                        "across %d rank(s), wall clock between barriers" % (2500 * copies, copies, world),
            "timings_rank0": {k: (round(v, 3) if isinstance(v, float) else v)
                              for k, v in driver.LAST_TIMINGS.items()}}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--sims", type=int, default=SIMS_PER_STEP, help="reactions per step per GPU")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-genotyping", action="store_true")
    ap.add_argument("--scaling-copies", type=int, default=40,
                    help="size of the strong-scaling VCF: 2500 distinct loci x this many copies x 32 samples "
                         "(40 = BASELINE config 5, 100k loci)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference_arm(args)

    import numpy as np
    import torch
    import torch.distributed as dist

    from sonics_b200 import engine as eng
    from sonics_b200 import sonics as drop_in

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        # NCCL logs to stdout (the version banner at NCCL_DEBUG=WARN / INFO): stdout carries ONE JSON line only
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    e = eng.Engine(local)
    dev = e.device
    params = eng.make_params()
    alleles, _, _ = drop_in.get_alleles(CFG2)
    n = args.sims
    table = e.build_table_from_arrays(params, [alleles], [0.0], uid=[0])
    out = {"lnL": torch.empty((1, n), dtype=torch.float64, device=dev),
           "group": torch.empty((1, n), dtype=torch.int16, device=dev)}
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    h_lnL = torch.empty((1, n), dtype=torch.float64, pin_memory=True)
    h_grp = torch.empty((1, n), dtype=torch.int16, pin_memory=True)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def dev_step(i):
        # every rank owns the simulation indices [rank*n, (rank+1)*n) of step i's seed
        return e.simulate(params, table, n, seed=20261019 + i, sim_begin=rank * n, out_offset=0, out=out)

    # end to end through the public host API, as a stream of jobs: genotype string in host memory -> table build +
    # H2D -> simulate -> lnL / group copied to pinned host memory.  The device-to-host copy of step i runs on its own
    # stream while step i + 1 computes (two result buffers on either side of PCIe); every step's copies lie inside
    # the timed region, which ends when the last one has landed.
    copy_stream = torch.cuda.Stream(device=dev)
    e2e_out = [out, {k: torch.empty_like(v) for k, v in out.items()}]
    e2e_host = [(h_lnL, h_grp), (torch.empty((1, n), dtype=torch.float64, pin_memory=True),
                                 torch.empty((1, n), dtype=torch.int16, pin_memory=True))]
    e2e_copied = [None, None]
    e2e_tables = []

    def e2e_step(i):
        k = i & 1
        al, _, _ = drop_in.get_alleles(CFG2)
        t = e.build_table_from_arrays(params, [al], [0.0], uid=[0])
        e2e_tables.append(t)
        if e2e_copied[k] is not None:  # the copy of step i - 2 has left this result buffer
            torch.cuda.current_stream().wait_event(e2e_copied[k])
        o = e.simulate(params, t, n, seed=777 + i, sim_begin=rank * n, out_offset=0, out=e2e_out[k])
        done = torch.cuda.Event()
        done.record()
        copy_stream.wait_event(done)
        with torch.cuda.stream(copy_stream):
            e2e_host[k][0].copy_(o["lnL"], non_blocking=True)
            e2e_host[k][1].copy_(o["group"], non_blocking=True)
            e2e_copied[k] = torch.cuda.Event()
            e2e_copied[k].record(copy_stream)
        return t

    def e2e_drain():
        copy_stream.synchronize()
        torch.cuda.synchronize()
        del e2e_tables[:-1]

    # ---- warm-up ----
    for i in range(max(3, args.warmup)):
        dev_step(-1 - i)
    e2e_step(-1)
    e2e_drain()
    barrier()

    # ---- device-resident timing: K steps, CUDA events on the launching stream, L2 flushed between steps ----
    clocks = ClockSampler(local)
    clocks.start()
    launches0 = e.launches
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    t_wall0 = time.perf_counter()
    for i in range(args.steps):
        flush.zero_()
        ev[i][0].record()
        dev_step(i)
        ev[i][1].record()
    barrier()
    t_wall = time.perf_counter() - t_wall0
    launches = e.launches - launches0
    clk = clocks.stop()
    dev_ms = sum(a.elapsed_time(b) for a, b in ev)

    # ---- end-to-end timing through the host API ----
    barrier()
    t0 = time.perf_counter()
    for i in range(args.steps):
        tbl = e2e_step(i)
    e2e_drain()
    barrier()
    e2e_s = time.perf_counter() - t0
    h2d = int(tbl.buf.numel())
    d2h = int(h_lnL.numel() * 8 + h_grp.numel() * 2)

    # sanity: the work was really done (sentinel fraction and group spread of the last step)
    lnl_host = e2e_host[(args.steps - 1) & 1][0].numpy()[0]
    frac_sentinel = float((lnl_host == -999999.0).mean())
    assert 0.05 < frac_sentinel < 0.5 and np.isfinite(lnl_host).all(), frac_sentinel

    times = torch.tensor([dev_ms, e2e_s * 1e3], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    dev_ms, e2e_ms = float(times[0]), float(times[1])

    # ---- other single-locus configurations and the readout arm (rank 0, device-resident, 1 warm-up + 2 timed) ----
    def timed_rate(spec, kw, nn, readout=False):
        import math  # noqa: F401
        from sonics_b200 import driver
        kind, (reads, nc) = driver.prefilter(("s", "b", ".", spec), {"one_allele_threshold": 45,
                                                                    "noise_threshold": 0.05}, {"monte_carlo": True})
        pr = eng.make_params(**kw)
        tb = e.build_table_from_arrays(pr, [reads], [float(np.float32(nc))])
        o = e.simulate(pr, tb, nn, seed=5, readout=readout)
        e.sync()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        flush.zero_()
        a.record()
        for k in range(2):
            e.simulate(pr, tb, nn, seed=6 + k, readout=readout, out=o)
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / 2
        return {"reactions_per_s": nn / ms * 1e3, "ms_per_launch": ms, "reactions_per_launch": nn,
                "window_bins": int(tb.max_window)}

    configs = None
    if rank == 0:
        import math
        cfg3 = ";".join("%d|%d" % (a, int(5 + 400 * math.exp(-0.5 * ((a - 28) / 6) ** 2) +
                                          300 * math.exp(-0.5 * ((a - 36) / 5) ** 2))) for a in range(10, 50))
        configs = {
            "cfg1": dict(timed_rate("8|5;9|113;10|89", {}, n // 2), what="README example, defaults"),
            "cfg3": dict(timed_rate(cfg3, dict(n_cycles=30, capture_cycle=16), n // 16),
                         what="40 input alleles, 30 cycles, capture at 16, automatic noise"),
            "cfg2_fast_samplers": None,
            "cfg2_with_readout": dict(timed_rate(CFG2, {}, n // 2, readout=True),
                                      what="cfg2 with the per-simulation sequencing readout, identity and r^2 "
                                           "(sonics.pyx:341-384) -- everything the reference's one_repeat() does"),
        }

    if rank == 0:
        # beside the exact number, never instead of it: the opt-in approximate sampler (sonics_set_samplers)
        e.set_samplers("fast", 100.0)
        try:
            configs["cfg2_fast_samplers"] = dict(
                timed_rate(CFG2, {}, n // 2),
                what="cfg2 with SONICS_SAMPLERS_FAST, min_npq = 100: binomial draws with n p q >= 100 from a normal / "
                     "Cornish-Fisher approximation (opt-in, not the reference's sampler; gates in "
                     "tests/test_gpu_fast_samplers.py); `value` above is the exact mode")
        finally:
            e.set_samplers("exact")

    # ---- genotyping: strong scaling of a cfg5-style VCF, file to file, across all ranks ----
    geno_scale = None
    if not args.no_genotyping:
        import shutil
        import tempfile
        workdir = None
        if rank == 0:
            workdir = tempfile.mkdtemp(prefix="sonics_bench_")
        if world > 1:
            box = [workdir]
            dist.broadcast_object_list(box, src=0)
            workdir = box[0]
        try:
            geno_scale = genotyping_file_to_file(rank, world, barrier, workdir, args.scaling_copies)
        finally:
            barrier()
            if rank == 0:
                shutil.rmtree(workdir, ignore_errors=True)

    if rank == 0:
        total = float(n) * world * args.steps
        value = total / (dev_ms * 1e-3)
        kernel_ms = dev_ms / args.steps
        peaks = {}
        pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(pk):
            peaks = json.load(open(pk))
        hbm_peak = peaks.get("hbm_gbs", 6650.0)
        achieved_gbs = n * BYTES_PER_REACTION / (kernel_ms * 1e-3) / 1e9
        sm_mhz = clk.get("sm_mhz") or peaks.get("sm_max_mhz", 1965.0)
        peak_issue = 148 * 4 * sm_mhz * 1e6  # warp instructions / s at the clock seen during the run
        imin = {}
        pi = os.path.join(ROOT, "profiles", "k1_imin.json")
        if os.path.exists(pi):
            imin = json.load(open(pi)).get("per_config", {})

        def algorithmic(rate, peak, name):
            i = imin.get(name, {}).get("I_min_thread_inst_per_reaction")
            if not i:
                return None
            return {"I_min_thread_inst_per_reaction": i, "frac": i * rate / (peak * 32),
                    "source": "profiles/k1_imin.json (tests/imin_audit.py)"}
        roof = {"bound": "alu_issue", "achieved": None, "peak": peak_issue, "unit": "warp-inst/s", "frac": None,
                "traffic": None, "peak_source": "148 SMs x 4 schedulers x SM clock sampled during the run"}
        cnt = load_inst_per_reaction()
        sha = kernel_sha()
        if cnt and cnt.get("kernel_sha") == sha:
            ach = cnt["warp_inst_per_reaction"] * n / (kernel_ms * 1e-3)
            roof.update({
                "achieved": ach, "frac": ach / peak_issue,
                "warp_inst_per_reaction": cnt["warp_inst_per_reaction"],
                "lanes_active_of_32": cnt.get("avg_active_threads"),
                # lane-weighted view: thread-instructions executed per reaction (ncu) against the thread-instruction
                # peak, and the same for the instructions one simulation needs when it runs alone (I_min,
                # tests/imin_audit.py: the oracle's draw counts priced with this kernel's SASS): reactions/s x I /
                # (peak_issue x 32)
                "executed_thread_inst": {
                    "per_reaction": cnt.get("thread_inst_per_reaction"),
                    "frac": (cnt.get("thread_inst_per_reaction", 0.0) * n / (kernel_ms * 1e-3)) / (peak_issue * 32)},
                "achieved_algorithmic": algorithmic(n / (kernel_ms * 1e-3), peak_issue, "cfg2"),
                "per_config": cnt.get("per_config"),
                "traffic": cnt.get("dram_bytes_per_reaction", 0.0) * n if cnt.get("dram_bytes_per_reaction") else None,
                "traffic_note": "DRAM bytes of every kernel of the step (k_sort_keys + radix sort + k_pcr_ls) from "
                                "ncu, per reaction x reactions of this step",
                "source": cnt.get("source"), "kernel_sha": sha})
        else:
            roof["note"] = ("profiles/k1_counters.json is %s: instructions per reaction unknown for this build "
                            "(kernel sources hash %s)" % ("missing" if not cnt else "stale (captured from %s)"
                                                          % cnt.get("kernel_sha"), sha))
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(3, args.warmup), "ms_per_step": kernel_ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32 samplers / int32 counts / f64 lnL", "data": "synthetic",
            "config": {"workload": "cfg2: single locus %s, default constants (24 cycles, capture at 13, 300000 start "
                                   "copies), fixed-count simulation without the per-simulation readout (genotyping "
                                   "replays it for the reported row only), %d reactions per step per GPU" % (CFG2, n),
                       "l2": "flushed between steps (256 MiB memset, untimed)", "seed": 20261019,
                       "kernel": "lockstep (sort by priors + 32 similar simulations per warp)"},
            "e2e": {"value": total / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h},
            "gpu_launches": int(launches),
            "clocks": clk,
            "roofline": roof,
            "hbm_roofline": {"bound": "hbm", "achieved": achieved_gbs, "peak": hbm_peak, "unit": "GB/s",
                             "frac": achieved_gbs / hbm_peak, "peak_source": "measured" if peaks else "fallback",
                             "note": "%d B/reaction of mandatory HBM traffic: not the bound" % BYTES_PER_REACTION},
            "wall_s_timed_region": t_wall,
            "sentinel_fraction": frac_sentinel,
            "configs": configs,
        }
        for name in ("cfg1", "cfg3"):
            if configs and name in configs:
                configs[name]["achieved_algorithmic"] = algorithmic(configs[name]["reactions_per_s"], peak_issue, name)
        geno = {}
        items = None
        if world == 1 and not args.no_genotyping:
            geno, items = genotyping_throughput(e, eng, drop_in)
        if geno_scale:
            geno["strong_scaling"] = geno_scale
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline()
            if items:
                geno["cpu_baseline"] = genotyping_cpu_baseline(items)
        if geno:
            line["genotyping"] = geno
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
