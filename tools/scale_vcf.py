#!/usr/bin/env python
"""Genotyping throughput on a large synthetic VCF, file to file, on 1..8 GPUs (BASELINE.json config 5).

    python tools/scale_vcf.py --loci 100000 --samples 32                       # 1 GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 \\
        --master-port 29700 tools/scale_vcf.py --loci 100000 --samples 32      # 8 GPUs

The VCF is generated first (untimed; in chunks over the host cores of all ranks, identical for every
world size), then `sonics -m <vcf>` runs through driver.run_sonics: per-rank byte-range ingest, GPU
repeat-until-pass loop, per-rank row writer, rank 0 joins the parts.  Timed from before the ingest to the
joined output file, between barriers, max over ranks.  Prints one JSON line on rank 0."""
import argparse
import hashlib
import json
import multiprocessing as mp
import os
import sys
import time

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))

N_CHUNKS = 64


def _gen_chunk(job):
    import synth_vcf
    path, c, lo, hi, samples, seed, cov, noisy = job
    loci = synth_vcf.make_genotypes(hi - lo, samples, seed + c, cov, noisy)
    synth_vcf.write_vcf("%s.c%02d" % (path, c), loci, samples, start=lo, header=(c == 0))
    return c


def generate(path, loci, samples, seed, cov, noisy, rank, world):
    bounds = [(c * loci // N_CHUNKS, (c + 1) * loci // N_CHUNKS) for c in range(N_CHUNKS)]
    jobs = [(path, c, lo, hi, samples, seed, cov, noisy) for c, (lo, hi) in enumerate(bounds) if c % world == rank]
    procs = max(1, min(len(jobs), (os.cpu_count() or 1) // world))
    with mp.get_context("fork").Pool(procs) as pool:
        pool.map(_gen_chunk, jobs)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--loci", type=int, default=100000)
    ap.add_argument("--samples", type=int, default=32)
    ap.add_argument("--cov", type=int, nargs=2, default=(2000, 10000))
    ap.add_argument("--noisy", type=float, default=0.1)
    ap.add_argument("--seed", type=int, default=20261019)
    ap.add_argument("--reps", type=int, default=1000)
    ap.add_argument("--dir", default="/tmp/sonics_scale")
    a = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    os.makedirs(a.dir, exist_ok=True)
    vcf = os.path.join(a.dir, "cfg5_%dx%d.vcf" % (a.loci, a.samples))

    t0 = time.time()
    generate(vcf, a.loci, a.samples, a.seed, tuple(a.cov), a.noisy, rank, world)  # before any CUDA / NCCL state
    import torch
    import torch.distributed as dist
    from sonics_b200 import driver
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG", "WARN")
        local = int(os.environ.get("LOCAL_RANK", rank))
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        dist.barrier()
    if rank == 0:
        with open(vcf, "wb") as fo:
            for c in range(N_CHUNKS):
                part = "%s.c%02d" % (vcf, c)
                with open(part, "rb") as fi:
                    fo.write(fi.read())
                os.unlink(part)
    t_gen = time.time() - t0
    args = driver.build_parser().parse_args([vcf, "-m", "-o", a.dir, "-n", "scale_out.txt", "-r", str(a.reps),
                                             "--seed", str(a.seed)])
    constants, ranges, options = driver.dicts_from_args(args)
    from sonics_b200 import sonics as drop_in
    drop_in.engine(int(os.environ.get("LOCAL_RANK", 0)))  # context + library load outside the timed region
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.time()
    driver.run_sonics(args.INPUT, constants, ranges, options, seed=a.seed, chunk=args.chunk)
    if world > 1:
        dist.barrier()
    dt = time.time() - t0
    if world > 1:
        t = torch.tensor([dt], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = float(t.item())
    if rank == 0:
        out = os.path.join(a.dir, "scale_out.txt")
        h = hashlib.sha256()
        rows = 0
        filt = {}
        with open(out, "rb") as f:
            for i, line in enumerate(f):
                h.update(line)
                if i:
                    rows += 1
                    k = line.split(b"\t")[7].decode()
                    filt[k] = filt.get(k, 0) + 1
        top = dict(sorted(filt.items(), key=lambda kv: -kv[1])[:6])
        print(json.dumps({"metric": "genotypes_per_s_file_to_file", "value": rows / dt, "unit": "genotypes/s",
                          "n_gpus": world, "rows": rows, "seconds": dt, "vcf_bytes": os.path.getsize(vcf),
                          "generate_seconds": t_gen, "timings_rank0": driver.LAST_TIMINGS,
                          "workload": "cfg5-style synthetic VCF %d loci x %d samples, coverage %d-%d, %.0f %% noisy, "
                                      "-r %d" % (a.loci, a.samples, a.cov[0], a.cov[1], 100 * a.noisy, a.reps),
                          "filter_hist_top": top, "output_sha256": h.hexdigest()}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
