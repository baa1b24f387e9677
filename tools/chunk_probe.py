#!/usr/bin/env python
"""cfg5-style VCF file to file on one GPU for several device-batch sizes (--chunk) and piece sizes.
usage: chunk_probe.py [copies=10] [chunk,chunk,...] [piece_mb,piece_mb,...]"""
import hashlib
import logging
import os
import sys
import tempfile
import time

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from sonics_b200 import driver  # noqa: E402

copies = int(sys.argv[1]) if len(sys.argv) > 1 else 10
chunks = [int(x) for x in (sys.argv[2] if len(sys.argv) > 2 else "65536,262144").split(",")]
pieces = [int(x) for x in (sys.argv[3] if len(sys.argv) > 3 else "8").split(",")]
logging.disable(logging.WARNING)
with tempfile.TemporaryDirectory() as d:
    vcf = os.path.join(d, "cfg5.vcf")
    n = bench.cfg5_subset_vcf(vcf, copies=copies)
    warm = os.path.join(d, "warm.vcf")
    bench.cfg5_subset_vcf(warm, distinct_loci=64, copies=1)
    args = driver.build_parser().parse_args(["-m", warm, "-o", d, "--seed", "1"])
    c, r, o = driver.dicts_from_args(args)
    driver.run_sonics(warm, c, r, o, seed=1)
    for pm in pieces:
        driver.PIECE_BYTES = pm << 20
        for ch in chunks:
            args = driver.build_parser().parse_args(["-m", vcf, "-o", d, "--seed", "20261019"])
            c, r, o = driver.dicts_from_args(args)
            t0 = time.perf_counter()
            driver.run_sonics(vcf, c, r, o, seed=20261019, chunk=ch)
            dt = time.perf_counter() - t0
            h = hashlib.sha256(open(os.path.join(d, "sonics_out.txt"), "rb").read()).hexdigest()[:12]
            print("piece %d MiB chunk %d: %.3f s  %.0f genotypes/s  sha %s  %s" % (
                pm, ch, dt, n / dt, h, {k: (round(v, 3) if isinstance(v, float) else v)
                                        for k, v in driver.LAST_TIMINGS.items()}), flush=True)
