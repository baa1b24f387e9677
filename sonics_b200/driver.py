"""Host driver: the reference's CLI script `sonics` (run_sonics / parse_vcf /
process_one_genotype / main, sonics:160-761) with the per-genotype map replaced by
one batched GPU call per shard.

What is kept byte-compatible: the flags and defaults (sonics:445-693), the two
output headers (sonics:174-207), the 0-/1-allele shortcuts and the auto-noise
rule of process_one_genotype (sonics:353-419), `--strict` handling in the VCF
parser (sonics:311-334) and the row format (via sonics_b200.sonics.format_row).

What changes: genotypes are gathered first, pre-filtered on the host, then run
through sonics_genotype_batch in chunks.  Under torchrun (one process per GPU)
the ranks share everything: each parses the lines of its byte range of the VCF,
simulates them and writes its rows; rank 0 joins the parts in rank order, so the
file is in input order (the reference appends from pool workers in completion
order, sonics:436).  With --gpus N in one process the list is split into
contiguous shards and the rows are written once by the parent.
"""
import argparse
import logging
import os
import re
import time

import numpy as np

from . import sonics as sonics_mod

VERSION = "0.3.0"  # the reference's VERSION file

DESC = ("SONiCS - Stutter mONte Carlo Simulation (B200-native engine). Models independent PCR reactions with "
        "reaction parameters drawn from weak uniform priors, scores each simulated PCR pool against the input "
        "readout with a log-likelihood, and calls the genotype with a Mann-Whitney U test and lnL ratios. "
        "Same interface as kzkedzierska/sonics; the Monte-Carlo engine runs on the GPU.")

HEADER_MC = ["sample", "block", "ref", "genotype", "identity", "identity_75th_percentile", "r_squared",
             "r_squared_75th_percentile", "lnL", "lnL_75th_percentile", "repeats\n"]
HEADER = ["sample", "block", "ref", "genotype", "identity", "r_squared", "lnL", "filter", "MWUtes_pval",
          "best_lnL_ratio", "quartile_lnL_ratio", "repeats", "additional_ratios\n"]


def build_parser():
    p = argparse.ArgumentParser(prog="sonics", description=DESC)
    p.add_argument("INPUT", type=str,
                   help="Either: allele composition - number of reads per allele, example: all1|reads;all2|reads "
                        "or path to a VCF file if run with --vcf_mode option")
    p.add_argument("-o", "--out_path", type=str, default=".", metavar="out_path")
    p.add_argument("-n", "--file_name", type=str, default="sonics_out.txt", metavar="file_name")
    p.add_argument("-p", "--processes", type=int, default=0, metavar="n_processes",
                   help="Kept for interface compatibility; the GPU engine batches genotypes instead of forking.")
    p.add_argument("-t", "--strict", type=int, default=1, metavar="n_strict")
    p.add_argument("-y", "--pcr_cycles", type=int, default=12, metavar="n_pcr_cycles")
    p.add_argument("-c", "--after_capture", type=int, default=12, metavar="n_after_capture")
    p.add_argument("-b", "--block", type=str, default="Block", metavar="block_id")
    p.add_argument("-a", "--name", type=str, default="sample", metavar="sample_name")
    p.add_argument("-r", "--repetitions", default=1000, type=int, metavar="n_repetitions")
    p.add_argument("-g", "--padjust", metavar="p_adjust_threshold", type=float, default=0.001)
    p.add_argument("-i", "--lnL_threshold", metavar="lnL_threshold", type=float, default=2.3)
    p.add_argument("-s", "--start_copies", default=300000, type=int, metavar="n_star_copies")
    p.add_argument("-l", "--noise_threshold", default=0.05, type=float, metavar="noise_threshold")
    p.add_argument("-f", "--floor", type=int, default=5, metavar="n_floor")
    p.add_argument("--min_sim", type=int, default=25, metavar="min_sim")
    p.add_argument("-j", "--one_allele", type=int, default=45, metavar="n_one_allele_threshold")
    p.add_argument("-e", "--efficiency", nargs=2, default=(0.001, 0.1), metavar=("MIN", "MAX"), type=float)
    p.add_argument("-d", "--down", nargs=2, metavar=("MIN", "MAX"), type=float, default=(0, 0.1))
    p.add_argument("-u", "--up", nargs=2, metavar=("MIN", "MAX"), type=float, default=(0, 0.1))
    p.add_argument("-k", "--capture", nargs=2, default=(-0.25, 0.25), type=float, metavar=("MIN", "MAX"))
    p.add_argument("--add_ratios", default="", type=str, metavar="ratios")
    p.add_argument("-m", "--vcf_mode", action="store_true", default=False)
    p.add_argument("--monte_carlo", action="store_true", default=False)
    p.add_argument("--random", action="store_true")
    p.add_argument("--save_report", action="store_true", default=False)
    p.add_argument("--up_preference", action="store_true")
    p.add_argument("-v", "--verbose", action="store_true")
    p.add_argument("--version", action="version", version="%(prog)s {}".format(VERSION))
    # extensions (not in the reference)
    p.add_argument("--seed", type=int, default=None,
                   help="Philox seed for a reproducible run (default: fresh entropy, like the reference).")
    p.add_argument("--gpus", type=int, default=1, help="GPUs of this box to shard the genotype list over.")
    p.add_argument("--chunk", type=int, default=131072, help="Genotypes per device batch.")
    p.add_argument("--fast_samplers", type=float, nargs="?", const=100.0, default=0.0, metavar="MIN_NPQ",
                   help="Opt-in approximate sampler (not in the reference): binomial draws with n*p*q >= MIN_NPQ "
                        "(default 100, at least 20) come from a normal / Cornish-Fisher approximation instead of the "
                        "exact rejection sampler. About 1.5x faster; passes the same KS / concordance gates, but "
                        "results differ draw by draw from the default mode.")
    return p


def dicts_from_args(args):
    """constants / ranges / sonics_run_options exactly as main() assembles them (sonics:711-751)."""
    capture_cycle = args.pcr_cycles + 1
    n_cycles = args.pcr_cycles + args.after_capture
    constants = {
        "random": args.random, "n_cycles": n_cycles, "capture_cycle": capture_cycle,
        "up_preference": args.up_preference, "floor": args.floor, "lnL_threshold": args.lnL_threshold,
        "start_copies": args.start_copies, "genotype": args.INPUT, "padjust": args.padjust,
        "noise_threshold": args.noise_threshold, "one_allele_threshold": args.one_allele}
    ranges = (args.down, args.up, args.capture, args.efficiency)
    options = {
        "repetitions": args.repetitions, "out_path": args.out_path, "strict": args.strict,
        "file_name": args.file_name, "vcf_mode": args.vcf_mode, "name": args.name, "block": args.block,
        "processes": args.processes, "verbose": bool(args.verbose), "save_report": args.save_report,
        "monte_carlo": args.monte_carlo, "add_ratios": args.add_ratios, "min_sim": args.min_sim}
    return constants, ranges, options


def parse_vcf(file_path, strict=1, rng=None):
    """lobSTR VCF -> [(sample, block, ref, 'allele|reads;...'), ...]   (sonics:246-343).

    Alleles are bp offsets from the reference; they become repeat counts offset / len(MOTIF) + REF.
    Partial repeats: strict 1 drops the allele, strict 2 drops the sample's genotype, strict 0 moves
    the support to one of the two neighbouring whole alleles at random.  Two reference defects are
    NOT reproduced: strict 0 indexes the support array by allele value
    (IndexError on the README's own example) and strict 2 forgets to advance the sample column."""
    rng = rng or np.random.default_rng()
    genotypes_list = []
    regexp_ref = re.compile(".*REF=([0-9.]*);.*")
    regexp_motif = re.compile(".*MOTIF=([TCGA]*);.*")
    samples = []
    with open(file_path, "r") as fh:
        for line in fh:
            if line.startswith("#"):
                if line.startswith("#CHROM"):
                    samples = line.strip("\n").split("\t")[9:]
                continue
            line_list = line.rstrip("\n").split("\t")
            block_name = line_list[0]
            info_field = line_list[7]
            m = regexp_ref.search(info_field)
            if m is None:
                raise Exception("There's something wrong with the vcf file, expecting REF in the INFO field "
                                "in the 8th column.")
            ref = int(m.group(1).split(".")[0])
            m = regexp_motif.search(info_field)
            if m is None:
                raise Exception("There's something wrong with the vcf file, expecting MOTIF in the INFO field "
                                "in the 8th column.")
            motif_length = len(m.group(1))
            all_reads_loc = line_list[8].split(":").index("ALLREADS")
            for k, sample in enumerate(samples):
                try:
                    gt_string = line_list[9 + k].split(":")[all_reads_loc]
                except IndexError:
                    logging.warning("Missing ALLREADS field in the input for block: %s, sample: %s", block_name, sample)
                    continue
                gt_list = gt_string.split(";")
                alleles = np.array([int(f.split("|")[0]) / motif_length + ref for f in gt_list])
                support = np.array([f.split("|")[1] for f in gt_list], dtype=int)
                diff = alleles % 1
                partial = diff != 0
                if partial.any():
                    if strict == 2:
                        logging.warning("Partial repetition of the motif in block: %s sample: %s. Excluding.",
                                        block_name, sample)
                        continue
                    elif strict == 0:
                        logging.warning("Partial repetition of the motif in block: %s, sample: %s. Adding support "
                                        "from partial allele to one of the closest allele on random.",
                                        block_name, sample)
                        whole = {int(a): i for i, a in enumerate(alleles) if not partial[i]}
                        extra_a, extra_s = [], []
                        for i in np.nonzero(partial)[0]:
                            target = int(np.floor(alleles[i])) + int(rng.integers(0, 2))
                            if target in whole:
                                support[whole[target]] += support[i]
                            else:
                                whole[target] = len(alleles) + len(extra_a)
                                extra_a.append(float(target))
                                extra_s.append(int(support[i]))
                            support[i] = 0
                        if extra_a:
                            alleles = np.concatenate([alleles, extra_a])
                            support = np.concatenate([support, extra_s])
                            order = np.argsort(alleles, kind="stable")
                            alleles, support = alleles[order], support[order]
                    else:
                        logging.warning("Partial repetition of the motif in block: %s sample: %s. Excluding allele.",
                                        block_name, sample)
                        support[partial] = 0
                genot = ";".join("%i|%i" % (a, s) for a, s in zip(alleles, support) if s > 0 and a > 0)
                genotypes_list.append((sample, block_name, ref, genot))
    return genotypes_list


def prefilter(input_tuple, constants, options):
    """The part of process_one_genotype() that runs before monte_carlo() (sonics:345-419).

    Returns ('skip', None) | ('row', row_string) | ('sim', (alleles, noise_coef))."""
    name, block, ref, genotype = input_tuple
    alleles, _, n_alleles = sonics_mod.get_alleles(genotype)
    total = sum(alleles)
    if n_alleles == 0:
        logging.warning("No alleles provided in input, skipping this genotype: %s for sample: %s.", block, name)
        return "skip", None
    if n_alleles == 1:
        if total > constants["one_allele_threshold"]:
            one_allele = alleles.nonzero()[0][0]
            if options["monte_carlo"]:
                # the reference builds this row but never writes it (sonics:364-374 vs :388-398)
                return "skip", None
            result = "\t".join(["{}/{}".format(one_allele, one_allele), ".", ".", ".", "no_simulations", ".", ".", ".",
                                "0", "."])
            logging.warning("Only one allele provided in input, with read support above the threshold. Including "
                            "this genotype: %s for sample: %s in the output. The simulation won't be performed and "
                            "likelihood won't be calculated.", block, name)
            return "row", result
        logging.warning("Only one allele provided in input, with read support below the threshold. Skipping this "
                        "genotype: %s for sample: %s in the output.", block, name)
        return "skip", None
    nz = alleles[alleles.nonzero()]
    frac = np.amin(nz) / total * np.count_nonzero(alleles)
    noise_frac = frac / (frac + 1)
    noise_coef = np.amin(nz) / total if noise_frac < constants["noise_threshold"] else 0
    return "sim", (alleles, float(noise_coef))


LAST_TIMINGS = {}  # seconds of the last run_sonics() in this process: ingest / simulate / write (tools/scale_vcf.py)


def shard_bounds(n, world, rank):
    """Contiguous shard [lo, hi) of n work items for `rank` of `world` (sizes differ by at most one)."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def run_batch_verdicts(allele_off, allele, reads, noise, constants, ranges, options, seed, device=0, chunk=131072,
                       uids=None):
    """Verdict array for a CSR batch of genotypes (device batches of `chunk`); uids key the RNG per genotype
    (default: the position in this batch)."""
    from . import engine as eng
    e = sonics_mod.engine(device)
    # --fast_samplers (main() exports it so that spawned per-GPU workers see it too); the dicts keep the reference's keys
    fast = float(os.environ.get("SONICS_FAST_SAMPLERS", "0") or 0.0)
    e.set_samplers("fast" if fast > 0 else "exact", fast if fast > 0 else 100.0)
    params = eng.params_from_dicts(constants, ranges, options)
    n = len(noise)
    uids = np.arange(n, dtype=np.uint32) if uids is None else np.ascontiguousarray(uids, dtype=np.uint32)
    parts = []
    # device batch: at most `chunk` genotypes, fewer when -r is large (workspace: 10 bytes per simulation, 22 with
    # --monte_carlo, within the free device memory)
    if n:
        chunk = e.batch_size_for(n, int(allele_off[n]), options["repetitions"], options.get("monte_carlo"), chunk)
    for c0 in range(0, n, chunk):
        c1 = min(n, c0 + chunk)
        off = allele_off[c0:c1 + 1] - allele_off[c0]
        a0, a1 = allele_off[c0], allele_off[c1]
        parts.append(e.genotype_batch(params, off, allele[a0:a1], reads[a0:a1], noise[c0:c1], options["repetitions"],
                                      seed, uid=uids[c0:c1]))
    return np.concatenate(parts) if parts else np.zeros(0, dtype=eng.VERDICT_DTYPE)


def _csr_slice(allele_off, allele, reads, noise, lo, hi):
    a0, a1 = allele_off[lo], allele_off[hi]
    return allele_off[lo:hi + 1] - a0, allele[a0:a1], reads[a0:a1], noise[lo:hi]


def shard_jobs(off, al, rd, nz, uids, gpus, common, chunk):
    """Argument tuples of _verdict_shard_worker: contiguous CSR shards, one per GPU, each with its own RNG keys."""
    jobs = []
    for r in range(gpus):
        lo, hi = shard_bounds(len(nz), gpus, r)
        jobs.append(_csr_slice(off, al, rd, nz, lo, hi) + common + (r, chunk, uids[lo:hi]))
    return jobs


def _verdict_shard_worker(args):
    off, al, rd, nz, constants, ranges, options, seed, device, chunk, uids = args
    return run_batch_verdicts(off, al, rd, nz, constants, ranges, options, seed, device, chunk, uids)


PIECE_BYTES = 8 << 20  # VCF bytes per pipeline piece (~140 k genotypes of a 32-sample lobSTR file)


def even_piece_bytes(size, world):
    """Piece size for `world` ranks: every rank gets the same number of pieces (a multiple of `world` pieces, at
    least 8 per rank so that the un-overlapped first parse and last write stay a small share, none above
    PIECE_BYTES, none below 1 MiB unless the file is that small).  With 22 pieces of 8 MiB on 8 ranks some ranks
    ran 3 pieces and some 2."""
    per_rank = max(8, -(-size // (world * PIECE_BYTES)))
    per_rank = max(1, min(per_rank, size // (world << 20)))  # keep pieces >= 1 MiB
    n = per_rank * world
    return max(1, -(-size // n))


def vcf_pipeline(feed_in, byte_range, line_base, out_path, header, constants, options, seed, compute,
                 piece_bytes=None, pieces=None, line_bases=None):
    """One process's share of VCF mode, pipelined: the file range is cut into pieces; while the GPU runs
    piece k (`compute`, a blocking native call that releases the GIL), one host thread parses piece k+1
    (sonics_vcf_parse_range) and another appends the rows of piece k-1 to out_path (sonics_rows_write_part).

    A genotype's RNG key is its cell in the (data line x sample) grid of the whole file: line_base = data
    lines before byte_range (hostio.lines_before), then every piece adds the lines it parsed -- so neither
    the pieces nor the ranks change the output.  With `pieces` (explicit [begin, end) list, not necessarily
    adjacent: the share of one rank of a multi-GPU run) every piece brings its own line base in `line_bases`.
    Returns (rows, simulated genotypes, timings dict); timings["row_bytes"] = bytes written per piece."""
    from concurrent.futures import ThreadPoolExecutor
    from . import hostio
    if pieces is None:
        lo, hi = byte_range
        end = os.path.getsize(feed_in) if hi < 0 else hi
        # pieces small enough that the un-overlapped first parse and last write are a small share of the range
        piece_bytes = piece_bytes or max(1 << 20, min(PIECE_BYTES, (end - lo) // 8))
        # a quarter-size first piece: its parse is the one the GPU waits for
        head = piece_bytes // 4 if end - lo > 2 * piece_bytes else 0
        cuts = ([lo] if head else []) + list(range(lo + head, end, piece_bytes)) + [end]
        pieces = [(cuts[i], cuts[i + 1]) for i in range(len(cuts) - 1)]
        if pieces and hi < 0:
            pieces[-1] = (pieces[-1][0], -1)
    tm = {"ingest": 0.0, "simulate": 0.0, "write": 0.0, "pieces": len(pieces), "row_bytes": []}
    n_rows = n_sim = 0

    def parse(rng):
        t = time.time()
        v = hostio.Vcf(feed_in, options["strict"], constants["one_allele_threshold"], constants["noise_threshold"],
                       options["monte_carlo"], seed, byte_range=rng)
        return v, time.time() - t

    def write(v, verdicts, first):
        t = time.time()
        before = 0 if first else os.path.getsize(out_path)
        hostio.write_rows(out_path, verdicts, v, options["monte_carlo"], options["add_ratios"], options["name"],
                          options["block"], header=header and first, append=not first)
        v.close()
        tm["row_bytes"].append(os.path.getsize(out_path) - before)
        return time.time() - t

    if not pieces:
        hostio.write_rows(out_path, np.zeros(0, dtype=hostio.VERDICT_DTYPE), None, options["monte_carlo"],
                          options["add_ratios"], options["name"], options["block"], header=header)
        return 0, 0, tm
    with ThreadPoolExecutor(1) as parser, ThreadPoolExecutor(1) as writer:
        nxt = parser.submit(parse, pieces[0])
        pending = None
        for k in range(len(pieces)):
            v, tp = nxt.result()
            tm["ingest"] += tp
            if k + 1 < len(pieces):
                nxt = parser.submit(parse, pieces[k + 1])
            if line_bases is not None:
                line_base = line_bases[k]
            uids = v.uids(line_base)
            line_base += v.n_blocks
            n_rows += v.n_rows
            n_sim += v.n_sim
            t = time.time()
            verdicts = compute(v.allele_off, v.allele, v.reads, v.noise_coef, uids)
            tm["simulate"] += time.time() - t
            if pending is not None:
                tm["write"] += pending.result()
            pending = writer.submit(write, v, verdicts, k == 0)
        tm["write"] += pending.result()
    return n_rows, n_sim, tm


def run_sonics(feed_in, constants, ranges, options, seed=None, gpus=1, chunk=131072):
    """run_sonics() of the reference (sonics:160-244) with the genotype loop batched on the GPU(s).

    VCF ingest, the pre-filters of process_one_genotype() and the output writer run natively
    (csrc/host_io.inl); parse_vcf() / prefilter() above are their Python mirrors (used with
    --save_report, which needs the per-genotype path, and by the tests that pin the native code)."""
    from . import hostio
    os.makedirs(options["out_path"], exist_ok=True)
    out_file_path = os.path.join(options["out_path"], options["file_name"])
    options["out_file_path"] = out_file_path
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if seed is None:
        seed = int.from_bytes(os.urandom(8), "little")
    if options["save_report"]:
        return _run_sonics_per_genotype(feed_in, constants, ranges, options, seed, out_file_path)
    if options["vcf_mode"] and world > 1:
        return _run_vcf_ranks(feed_in, constants, ranges, options, seed, chunk, out_file_path, rank, world)
    if options["vcf_mode"] and gpus <= 1:
        t0 = time.time()
        n_rows, n_sim, tm = vcf_pipeline(
            feed_in, (0, -1), 0, out_file_path, True, constants, options, seed,
            lambda off, al, rd, nz, uids: run_batch_verdicts(off, al, rd, nz, constants, ranges, options, seed, 0,
                                                             chunk, uids))
        LAST_TIMINGS.clear()
        tm.pop("row_bytes", None)
        LAST_TIMINGS.update(tm, join=0.0, genotypes=n_sim, wall=time.time() - t0)
        if n_rows == 0:
            raise Exception("No valid genotype after parsing VCF file.")
        return
    vcf = None
    if options["vcf_mode"]:  # --gpus N in one process: parse once, contiguous shards to N worker processes
        vcf = hostio.Vcf(feed_in, options["strict"], constants["one_allele_threshold"], constants["noise_threshold"],
                         options["monte_carlo"], seed)
        if vcf.n_rows == 0:
            raise Exception("No valid genotype after parsing VCF file.")
        off, al, rd, nz, uids = vcf.allele_off, vcf.allele, vcf.reads, vcf.noise_coef, vcf.uids(0)
    else:
        if rank != 0:
            return
        kind, payload = prefilter((options["name"], options["block"], ".", feed_in), constants, options)
        if kind == "skip":
            hostio.write_rows(out_file_path, np.zeros(0, dtype=hostio.VERDICT_DTYPE), None, options["monte_carlo"],
                              options["add_ratios"], options["name"], options["block"])
            return
        if kind == "row":
            with open(out_file_path, "w+") as f:
                f.write("\t".join(HEADER_MC if options["monte_carlo"] else HEADER))
                f.write("{}\t{}\t{}\t{}\n".format(options["name"], options["block"], ".", payload))
            return
        from . import engine as eng
        off, al, rd = eng.to_csr([payload[0]])
        nz = np.array([payload[1]], dtype=np.float32)
        uids = np.zeros(1, dtype=np.uint32)
    n = len(nz)
    if gpus > 1:
        import multiprocessing as mp
        jobs = shard_jobs(off, al, rd, nz, uids, gpus, (constants, ranges, options, seed), chunk)
        with mp.get_context("spawn").Pool(gpus) as pool:
            verdicts = np.concatenate(pool.map(_verdict_shard_worker, jobs))
    else:
        verdicts = run_batch_verdicts(off, al, rd, nz, constants, ranges, options, seed, 0, chunk, uids)
    hostio.write_rows(out_file_path, verdicts, vcf, options["monte_carlo"], options["add_ratios"], options["name"],
                      options["block"])


def _run_vcf_ranks(feed_in, constants, ranges, options, seed, chunk, out_file_path, rank, world, compute=None,
                   piece_bytes=None):
    """VCF mode under torchrun, one process per GPU: the ranks share ingest, simulation and output.

    The file is cut into byte pieces dealt round-robin (piece k -> rank k mod world: genotypes that need 100 or
    1000 simulations cluster by locus, so interleaving evens the load where contiguous shares did not).  Every
    rank counts the data lines of its own pieces (a memchr scan) and one all_gather of those counts gives every
    piece its line base -- the RNG key of a genotype stays its cell in the (data line x sample) grid of the whole
    file.  Rank r runs vcf_pipeline() over its pieces into <out>.part<r>; a second all_gather (bytes written per
    piece) gives every piece its offset in the output file and each rank copies its own pieces there.  No
    collective touches the data path.  `compute` replaces the GPU call in the CPU (gloo) test."""
    import torch.distributed as dist
    from . import hostio
    if not dist.is_initialized():
        import torch
        if torch.cuda.is_available():
            local = int(os.environ.get("LOCAL_RANK", rank))
            torch.cuda.set_device(local)
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        else:
            dist.init_process_group("gloo")
    t0 = time.time()
    local = int(os.environ.get("LOCAL_RANK", rank))
    compute = compute or (lambda off, al, rd, nz, uids: run_batch_verdicts(
        off, al, rd, nz, constants, ranges, options, seed, local, chunk, uids))
    size = os.path.getsize(feed_in)
    if piece_bytes:
        all_pieces = hostio.pieces_of(feed_in, piece_bytes)
    else:
        # every rank starts on a quarter-size piece (its parse is the one the GPU waits for); the rest of the file
        # in equal pieces, the same number for every rank
        head = even_piece_bytes(size, world) // 4
        all_pieces = hostio.pieces_of(feed_in, even_piece_bytes(max(0, size - world * head), world), world, head)
    mine = list(range(rank, len(all_pieces), world))
    gathered = [None] * world
    dist.all_gather_object(gathered, [hostio.count_lines(feed_in, *all_pieces[i]) for i in mine])
    n_lines = [0] * len(all_pieces)
    for r in range(world):
        for j, i in enumerate(range(r, len(all_pieces), world)):
            n_lines[i] = gathered[r][j]
    bases = np.concatenate([[0], np.cumsum(n_lines)]).astype(np.int64)
    part = "%s.part%d" % (out_file_path, rank)
    n_rows, n_sim, tm = vcf_pipeline(feed_in, None, 0, part, rank == 0, constants, options, seed, compute,
                                     pieces=[all_pieces[i] for i in mine], line_bases=[int(bases[i]) for i in mine])
    t1 = time.time()
    # join: every rank copies its pieces to their offsets of the output file (parallel; rank 0 sizes the file)
    dist.all_gather_object(gathered, (n_rows, tm["row_bytes"] if mine else [os.path.getsize(part)]))
    sizes = [0] * max(1, len(all_pieces))
    for r in range(world):
        for j, i in enumerate(range(r, len(all_pieces), world)):
            sizes[i] = gathered[r][1][j]
    if not all_pieces:
        sizes = [gathered[0][1][0]]  # header only, written by rank 0
    offs = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
    # every rank copies its pieces to their offsets of the output file (writes beyond the end extend it, so nobody
    # waits for the file to be sized; rank 0 cuts it to its final length -- a longer file of an earlier run -- last)
    fd_out = os.open(out_file_path, os.O_RDWR | os.O_CREAT, 0o644)
    fd_in = os.open(part, os.O_RDONLY)
    try:
        src = 0
        for j, i in enumerate(mine if all_pieces else ([0] if rank == 0 else [])):
            dst, left = int(offs[i]), int(sizes[i])
            while left > 0:
                try:
                    n = os.copy_file_range(fd_in, fd_out, left, src, dst)  # in-kernel copy, same file system
                except (OSError, AttributeError):
                    buf = os.pread(fd_in, min(left, 8 << 20), src)
                    n = os.pwrite(fd_out, buf, dst) if buf else 0
                if n <= 0:
                    break
                src, dst, left = src + n, dst + n, left - n
    finally:
        os.close(fd_in)
        os.close(fd_out)
    os.unlink(part)
    dist.barrier()
    if rank == 0:
        os.truncate(out_file_path, int(offs[-1]))
    LAST_TIMINGS.clear()
    tm.pop("row_bytes", None)
    LAST_TIMINGS.update(tm, join=time.time() - t1, genotypes=n_sim, wall=time.time() - t0, pieces_total=len(all_pieces))
    if sum(g[0] for g in gathered) == 0:
        raise Exception("No valid genotype after parsing VCF file.")


def _run_sonics_per_genotype(feed_in, constants, ranges, options, seed, out_file_path):
    """--save_report: every simulation record is written (sonics.pyx:110-113, 201-206), one genotype at a time."""
    if options["vcf_mode"]:
        parsed = parse_vcf(feed_in, options["strict"])
        if not parsed:
            raise Exception("No valid genotype after parsing VCF file.")
    else:
        parsed = [(options["name"], options["block"], ".", feed_in)]
    sonics_mod.seed(seed)
    with open(out_file_path, "w+") as out_file:
        out_file.write("\t".join(HEADER_MC if options["monte_carlo"] else HEADER))
        for t in parsed:
            kind, payload = prefilter(t, constants, options)
            if kind == "skip":
                continue
            if kind == "sim":
                c = dict(constants, alleles=payload[0], noise_coef=payload[1])
                o = dict(options, name=t[0], block=t[1])
                payload = sonics_mod.monte_carlo(options["repetitions"], c, ranges, o)
            logging.info(payload)
            out_file.write("{}\t{}\t{}\t{}\n".format(t[0], t[1], t[2], payload))


def main(argv=None):
    args = build_parser().parse_args(argv)
    logging.basicConfig(level=logging.DEBUG if args.verbose else logging.INFO,
                        format="%(asctime)s %(levelname)s %(message)s")
    logging.debug("SONiCS run with the following options:")
    logging.debug(args)
    constants, ranges, options = dicts_from_args(args)
    if args.fast_samplers:
        os.environ["SONICS_FAST_SAMPLERS"] = repr(float(args.fast_samplers))
    else:
        os.environ.pop("SONICS_FAST_SAMPLERS", None)
    run_sonics(args.INPUT, constants, ranges, options, seed=args.seed, gpus=args.gpus, chunk=args.chunk)


if __name__ == "__main__":
    main()
