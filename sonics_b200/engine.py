"""Host-side engine: owns the C-ABI context and the PyTorch device buffers.

PyTorch is plumbing only (device memory + streams); every computation happens in
the CUDA library behind include/sonics_b200.h.
"""
import ctypes
import os

import numpy as np

from . import _abi
from ._abi import SonicsParams, SonicsVerdict, check

MAX_ALLELE = 500

VERDICT_DTYPE = np.dtype([
    ("allele_lo", "<i4"), ("allele_hi", "<i4"), ("filter", "<i4"), ("run_reps", "<i4"), ("min_sim", "<i4"),
    ("best_sim", "<i4"), ("n_groups", "<i4"), ("p_clamped", "<i4"), ("lnL", "<f8"), ("identity", "<f8"),
    ("r_squared", "<f8"), ("high_pval", "<f8"), ("best_lnL_ratio", "<f8"), ("percentile_lnL_ratio", "<f8"),
    ("identity_q75", "<f8"), ("r_squared_q75", "<f8"), ("lnL_q75", "<f8"), ("add_ratio", "<f8", (8,))])
assert VERDICT_DTYPE.itemsize == ctypes.sizeof(SonicsVerdict)


def make_params(n_cycles=24, capture_cycle=13, floor=5, random=False, up_preference=False, start_copies=300000,
                min_sim=25, monte_carlo=False, down=(0, 0.1), up=(0, 0.1), capture=(-0.25, 0.25),
                efficiency=(0.001, 0.1), padjust=0.001, lnL_threshold=2.3, add_ratios=""):
    """SonicsParams from the reference's CLI defaults (sonics:445-693, 711-751)."""
    p = SonicsParams()
    p.n_cycles = int(n_cycles)
    p.capture_cycle = int(capture_cycle)
    p.floor_opt = int(floor)
    p.random_start = int(bool(random))
    p.up_preference = int(bool(up_preference))
    p.start_copies = int(start_copies)
    p.min_sim = int(min_sim)
    p.monte_carlo = int(bool(monte_carlo))
    p.down[0], p.down[1] = float(down[0]), float(down[1])
    p.up[0], p.up[1] = float(up[0]), float(up[1])
    p.capture[0], p.capture[1] = float(capture[0]), float(capture[1])
    p.efficiency[0], p.efficiency[1] = float(efficiency[0]), float(efficiency[1])
    p.padjust = float(padjust)
    p.lnL_threshold = float(lnL_threshold)
    qs = [float(x) for x in add_ratios.split(";")] if add_ratios else []  # sonics.pyx:232-235
    if len(qs) > 8:
        raise ValueError("at most 8 --add_ratios percentiles are supported")
    p.n_add_ratios = len(qs)
    for i, q in enumerate(qs):
        p.add_ratios[i] = q
    return p


def params_from_dicts(constants, ranges, options=None):
    """The reference's `constants` / `ranges` / `options` dicts (sonics:714-751) -> SonicsParams."""
    options = options or {}
    return make_params(
        n_cycles=constants["n_cycles"], capture_cycle=constants["capture_cycle"], floor=constants["floor"],
        random=constants["random"], up_preference=constants["up_preference"],
        start_copies=constants["start_copies"], min_sim=options.get("min_sim", 25),
        monte_carlo=options.get("monte_carlo", False), down=ranges[0], up=ranges[1], capture=ranges[2],
        efficiency=ranges[3], padjust=constants.get("padjust", 0.001),
        lnL_threshold=constants.get("lnL_threshold", 2.3), add_ratios=options.get("add_ratios", ""))


def to_csr(genotypes):
    """List of reads-per-repeat-count arrays (get_alleles()[0]) -> CSR (allele_off, allele, reads)."""
    off = np.zeros(len(genotypes) + 1, dtype=np.int32)
    al, rd = [], []
    for g, arr in enumerate(genotypes):
        arr = np.asarray(arr)
        nz = np.nonzero(arr)[0]
        al.append(nz.astype(np.int32))
        rd.append(arr[nz].astype(np.int32))
        off[g + 1] = off[g] + len(nz)
    return off, (np.concatenate(al) if al else np.zeros(0, np.int32)), (np.concatenate(rd) if rd else np.zeros(0, np.int32))


TRACE_DTYPE = np.dtype([("kind", "<i4"), ("n", "<i4"), ("p", "<f4"), ("x", "<i4")])  # SonicsTraceEntry


class Table:
    """A genotype table resident in device memory (torch uint8 tensor owned here)."""

    def __init__(self, buf, n_gt, allele_off, allele, reads, max_window, owner=None):
        self._owner = owner  # the Engine whose ctx holds the host-side record of this table
        self.buf = buf
        self.n_gt = n_gt
        self.allele_off = allele_off
        self.allele = allele
        self.reads = reads
        self.max_window = max_window

    def __del__(self):
        try:
            e = self._owner
            if e is not None and getattr(e, "ctx", None):
                e.lib.sonics_table_release(e.ctx, self.buf.data_ptr())
        except Exception:
            pass

    def group_label(self, g, gid):
        """u16 group id -> 'lo/hi' label (sonics.pyx:329-332)."""
        a0, a1 = self.allele_off[g], self.allele_off[g + 1]
        A = a1 - a0
        return "%d/%d" % (self.allele[a0 + gid // A], self.allele[a0 + gid % A])


class Engine:
    def __init__(self, device=0):
        import torch  # device memory + streams only

        if not torch.cuda.is_available():
            raise RuntimeError("sonics_b200: no CUDA device visible; this framework has no CPU fallback")
        self.torch = torch
        self.lib = _abi.load()
        self.device_index = int(device)
        self.device = torch.device("cuda", self.device_index)
        torch.cuda.set_device(self.device)
        torch.zeros(1, device=self.device)  # make sure the primary context exists
        ctx = ctypes.c_void_p()
        check(self.lib.sonics_create(self.device_index, ctypes.byref(ctx)))
        self.ctx = ctx
        # mirrors the library's SONICS_K1 measurement switch (scratch sizing differs between the two forms of K1)
        self.kernel = {"persistent": _abi.KERNEL_PERSISTENT, "lockstep": _abi.KERNEL_LOCKSTEP}.get(
            os.environ.get("SONICS_K1", ""), _abi.KERNEL_AUTO)
        self._pools_wanted = False

    def set_kernel(self, kernel):
        """Which form of K1 simulate() launches: 'auto' (lockstep unless tracing), 'persistent', 'lockstep'."""
        k = {"auto": _abi.KERNEL_AUTO, "persistent": _abi.KERNEL_PERSISTENT, "lockstep": _abi.KERNEL_LOCKSTEP}[kernel]
        check(self.lib.sonics_set_kernel(self.ctx, k))
        self.kernel = k

    def set_samplers(self, mode="exact", min_npq=100.0):
        """'exact' (default): inversion / BTRS.  'fast': binomial draws with n p q >= min_npq from the normal /
        Cornish-Fisher approximation (sonics_set_samplers; opt-in, not the reference's sampler)."""
        check(self.lib.sonics_set_samplers(self.ctx, {"exact": 0, "fast": 1}[mode], float(min_npq)))

    def close(self):
        if getattr(self, "ctx", None):
            self.lib.sonics_destroy(self.ctx)
            self.ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- helpers ---------------------------------------------------------
    def _stream(self):
        return ctypes.c_void_p(self.torch.cuda.current_stream(self.device).cuda_stream)

    def sync(self):
        check(self.lib.sonics_sync(self.ctx, self._stream()))

    @property
    def launches(self):
        return int(self.lib.sonics_launch_count(self.ctx))

    # -- test hooks -------------------------------------------------------
    def debug_philox(self, seed, sim, uid, stream_id, n_blocks):
        out = self.torch.empty(4 * n_blocks, dtype=self.torch.int32, device=self.device)
        check(self.lib.sonics_debug_philox(self.ctx, seed, sim, uid, stream_id, n_blocks, out.data_ptr()))
        return out.cpu().numpy().view(np.uint32)

    def debug_pairs(self, seed, sim, uid, n_pairs):
        out = self.torch.empty(2 * n_pairs, dtype=self.torch.int32, device=self.device)
        check(self.lib.sonics_debug_pairs(self.ctx, seed, sim, uid, n_pairs, out.data_ptr()))
        return out.cpu().numpy().view(np.uint32)

    def simulate_traced(self, params, table, sim_count, seed, trace_cap=4096, sim_begin=0):
        """simulate() through the traced instantiation of K1a (sonics_debug_trace).

        Returns (out, trace, trace_len, pools): out as simulate(simpar=True) returns it (host copies); trace
        [n_gt, sim_count, trace_cap] records (TRACE_DTYPE: kind, n, p, x); trace_len [n_gt, sim_count];
        pools [n_gt, sim_count, max_window] int32, the final pools in window coordinates."""
        t = self.torch
        units = table.n_gt * int(sim_count)
        trace = t.zeros(units * trace_cap * TRACE_DTYPE.itemsize, dtype=t.uint8, device=self.device)
        tlen = t.zeros(units, dtype=t.int32, device=self.device)
        check(self.lib.sonics_debug_trace(self.ctx, trace.data_ptr(), int(trace_cap), tlen.data_ptr()))
        self._pools_wanted = True  # the traced (persistent) form leaves the final pools in the scratch
        try:
            out = self.simulate(params, table, sim_count, seed, sim_begin=sim_begin, simpar=True, out_offset=0)
            self.sync()
        finally:
            self._pools_wanted = False
            check(self.lib.sonics_debug_trace(self.ctx, None, 0, None))
        W = table.max_window
        raw = self._scratch_buf[256:256 + ((units + 31) // 32) * W * 128].cpu().numpy().view(np.int32)
        pools = raw.reshape(-1, W, 32).transpose(0, 2, 1).reshape(-1, W)[:units]
        host = {k: v.cpu().numpy() for k, v in out.items()}
        return (host, trace.cpu().numpy().view(TRACE_DTYPE).reshape(table.n_gt, sim_count, trace_cap),
                tlen.cpu().numpy().reshape(table.n_gt, sim_count), pools.reshape(table.n_gt, sim_count, W))

    def debug_sample(self, kind, n, p, count, seed):
        out = self.torch.empty(count, dtype=self.torch.int32, device=self.device)
        check(self.lib.sonics_debug_sample(self.ctx, {"binomial": 0, "poisson": 1, "binomial_fast": 2}[kind], int(n), float(p), int(count),
                                           int(seed), out.data_ptr()))
        return out.cpu().numpy()

    # -- pymc_extracted.mvhyperg ------------------------------------------
    def mvhyperg(self, x, color):
        """mvhyperg(x, color) like the f2py routine; accepts one row or a 2-D batch of rows."""
        x = np.ascontiguousarray(x, dtype=np.int64)
        color = np.ascontiguousarray(color, dtype=np.int64)
        single = x.ndim == 1
        x2, c2 = np.atleast_2d(x), np.atleast_2d(color)
        n, k = x2.shape
        out = np.empty(n, dtype=np.float64)
        scratch = self.torch.empty(n * (2 * k * 8 + 8), dtype=self.torch.uint8, device=self.device)
        check(self.lib.sonics_mvhyperg(self.ctx, x2.ctypes.data, c2.ctypes.data, k, n, out.ctypes.data,
                                       scratch.data_ptr(), scratch.numel()))
        return float(out[0]) if single else out

    # -- genotype table -----------------------------------------------------
    def build_table(self, params, allele_off, allele, reads, noise_coef=None, uid=None):
        allele_off = np.ascontiguousarray(allele_off, dtype=np.int32)
        allele = np.ascontiguousarray(allele, dtype=np.int32)
        reads = np.ascontiguousarray(reads, dtype=np.int32)
        n_gt = len(allele_off) - 1
        nc = None if noise_coef is None else np.ascontiguousarray(noise_coef, dtype=np.float32)
        ui = None if uid is None else np.ascontiguousarray(uid, dtype=np.uint32)
        nbytes = self.lib.sonics_table_bytes(n_gt, int(allele_off[-1]))
        buf = self.torch.empty(max(nbytes, 16), dtype=self.torch.uint8, device=self.device)
        check(self.lib.sonics_table_build(
            self.ctx, ctypes.byref(params), n_gt, allele_off.ctypes.data, allele.ctypes.data, reads.ctypes.data,
            None if nc is None else nc.ctypes.data, None if ui is None else ui.ctypes.data, buf.data_ptr(),
            buf.numel(), self._stream()))
        return Table(buf, n_gt, allele_off, allele, reads, self.lib.sonics_table_max_window(self.ctx), owner=self)

    def build_table_from_arrays(self, params, genotypes, noise_coef=None, uid=None):
        off, al, rd = to_csr(genotypes)
        return self.build_table(params, off, al, rd, noise_coef, uid)

    # -- K1 -------------------------------------------------------------------
    def simulate(self, params, table, sim_count, seed, sim_begin=0, readout=False, simpar=False, out=None,
                 active=None, out_offset=None):
        """Run simulations j in [sim_begin, sim_begin+sim_count) for every (active) genotype.

        Results land at column out_offset + (j - sim_begin); out_offset defaults to sim_begin (cumulative
        per-genotype arrays).  Returns a dict of device tensors shaped [n_gt, stride]: lnL (f64), group
        (i16 view of the u16 ids) and, with readout=True, identity (f64) / rsq (f32); with simpar=True,
        simpar [n_gt, stride, 4] (f64: down, up, capture, efficiency)."""
        t = self.torch
        if out_offset is None:
            out_offset = sim_begin
        stride = out_offset + sim_count
        if out is None:
            out = {"lnL": t.empty((table.n_gt, stride), dtype=t.float64, device=self.device),
                   "group": t.empty((table.n_gt, stride), dtype=t.int16, device=self.device)}
            if readout:
                out["identity"] = t.empty((table.n_gt, stride), dtype=t.float64, device=self.device)
                out["rsq"] = t.empty((table.n_gt, stride), dtype=t.float32, device=self.device)
            if simpar:
                out["simpar"] = t.empty((table.n_gt, stride, 4), dtype=t.float64, device=self.device)
        stride = out["lnL"].shape[1]
        act_ptr, n_act = None, table.n_gt
        if active is not None:
            act_ptr, n_act = active.data_ptr(), int(active.numel())
        scratch = self._scratch(n_act * int(sim_count), table.max_window)
        check(self.lib.sonics_simulate(
            self.ctx, ctypes.byref(params), table.buf.data_ptr(), table.n_gt, act_ptr, n_act, int(sim_begin),
            int(sim_count), int(seed) & 0xFFFFFFFFFFFFFFFF, int(stride), int(out_offset), out["lnL"].data_ptr(),
            out["group"].data_ptr(), out["identity"].data_ptr() if "identity" in out else None,
            out["rsq"].data_ptr() if "rsq" in out else None, out["simpar"].data_ptr() if "simpar" in out else None,
            scratch.data_ptr(), scratch.numel(), self._stream()))
        return out

    SCRATCH_CAP = 2 << 30  # bytes; larger jobs run in chunks

    def _scratch(self, n_units, max_window):
        persistent = self._pools_wanted or self.kernel == _abi.KERNEL_PERSISTENT
        k = _abi.KERNEL_PERSISTENT if persistent else _abi.KERNEL_LOCKSTEP
        need = min(int(self.lib.sonics_scratch_bytes_kernel(int(n_units), int(max_window), k)), self.SCRATCH_CAP)
        # never below what either form needs for its smallest chunk (a counter / trace hook switches forms)
        for kk in (_abi.KERNEL_PERSISTENT, _abi.KERNEL_LOCKSTEP):
            need = max(need, int(self.lib.sonics_scratch_bytes_kernel(32, int(max_window), kk)))
        return self._scratch_bytes(need)

    def simulate_with_readout_dump(self, params, table, sim_count, seed, sim_begin=0):
        """simulate(readout=True) that also returns the simulated sequencing readouts (sonics_debug_readout):
        (out, readout) with readout [n_gt, sim_count, max_window] int32 in window coordinates (host)."""
        t = self.torch
        units = table.n_gt * int(sim_count)
        W = table.max_window
        dump = t.zeros(units * W, dtype=t.int32, device=self.device)
        check(self.lib.sonics_debug_readout(self.ctx, dump.data_ptr()))
        try:
            out = self.simulate(params, table, sim_count, seed, sim_begin=sim_begin, readout=True, out_offset=0)
            self.sync()
        finally:
            check(self.lib.sonics_debug_readout(self.ctx, None))
        return out, dump.cpu().numpy().reshape(table.n_gt, int(sim_count), W)

    def _scratch_bytes(self, need):
        s = getattr(self, "_scratch_buf", None)
        if s is None or s.numel() < need:
            self._scratch_buf = None
            s = self.torch.empty(need, dtype=self.torch.uint8, device=self.device)
            self._scratch_buf = s
        return s

    # -- K2/K3 ------------------------------------------------------------------
    def evaluate(self, params, table, out, n_run, final_round=True, active=None, verdict=None):
        """Statistics of sonics.pyx:136-252 over the first n_run simulations of every (active) genotype.

        `out` is the dict simulate() returns (or any dict with device tensors lnL [n_gt, stride] f64 and
        group [n_gt, stride] i16).  Returns a device uint8 tensor holding SonicsVerdict[n_gt]."""
        t = self.torch
        if verdict is None:
            verdict = t.zeros(table.n_gt * VERDICT_DTYPE.itemsize, dtype=t.uint8, device=self.device)
        act_ptr, n_act = None, table.n_gt
        if active is not None:
            act_ptr, n_act = active.data_ptr(), int(active.numel())
        max_a = int(np.max(np.diff(table.allele_off)))
        sbytes = min(int(self.lib.sonics_stats_scratch_bytes_batch(n_act, int(n_run), max_a, int(params.random_start))),
                     max(self.SCRATCH_CAP * 4, int(self.lib.sonics_stats_scratch_bytes(int(n_run), max_a,
                                                                                      int(params.random_start)))))
        sptr, snum = None, 0
        if sbytes:
            sc = self._scratch_bytes(sbytes)
            sptr, snum = sc.data_ptr(), sc.numel()
        check(self.lib.sonics_evaluate(
            self.ctx, ctypes.byref(params), table.buf.data_ptr(), table.n_gt, act_ptr, n_act, int(n_run),
            int(out["lnL"].shape[1]), out["lnL"].data_ptr(), out["group"].data_ptr(),
            out["identity"].data_ptr() if "identity" in out else None,
            out["rsq"].data_ptr() if "rsq" in out else None, int(bool(final_round)), verdict.data_ptr(),
            sptr, snum, self._stream()))
        return verdict

    def replay_best(self, params, table, seed, verdict):
        scratch = self._scratch(table.n_gt, table.max_window)
        check(self.lib.sonics_replay_best(self.ctx, ctypes.byref(params), table.buf.data_ptr(), table.n_gt, None,
                                          table.n_gt, int(seed) & 0xFFFFFFFFFFFFFFFF, verdict.data_ptr(),
                                          scratch.data_ptr(), scratch.numel(), self._stream()))
        return verdict

    def verdicts_to_host(self, verdict):
        self.sync()
        return verdict.cpu().numpy().view(VERDICT_DTYPE)

    # -- the whole repeat-until-pass loop ---------------------------------------
    def genotype_batch(self, params, allele_off, allele, reads, noise_coef, max_reps, seed, uid=None):
        """monte_carlo() for a batch of genotypes (host arrays in, host verdicts out)."""
        allele_off = np.ascontiguousarray(allele_off, dtype=np.int32)
        allele = np.ascontiguousarray(allele, dtype=np.int32)
        reads = np.ascontiguousarray(reads, dtype=np.int32)
        n_gt = len(allele_off) - 1
        nc = None if noise_coef is None else np.ascontiguousarray(noise_coef, dtype=np.float32)
        ui = None if uid is None else np.ascontiguousarray(uid, dtype=np.uint32)
        nbytes = self.lib.sonics_genotype_work_bytes_for(ctypes.byref(params), n_gt, allele_off.ctypes.data,
                                                         int(max_reps))
        work = self._workspace(nbytes)
        out = np.zeros(n_gt, dtype=VERDICT_DTYPE)
        check(self.lib.sonics_genotype_batch(
            self.ctx, ctypes.byref(params), n_gt, allele_off.ctypes.data, allele.ctypes.data, reads.ctypes.data,
            None if nc is None else nc.ctypes.data, None if ui is None else ui.ctypes.data, int(max_reps),
            int(seed) & 0xFFFFFFFFFFFFFFFF, out.ctypes.data, work.data_ptr(), work.numel(), self._stream()))
        return out

    def batch_size_for(self, n_gt, n_alleles_total, max_reps, monte_carlo, want, fraction=0.6):
        """Largest number of genotypes <= want whose genotype_batch workspace fits `fraction` of the free device
        memory (a VCF with a large -r is run in several device batches instead of failing to allocate)."""
        free, _ = self.torch.cuda.mem_get_info(self.device)
        budget = int(free * fraction) + (self._work.numel() if getattr(self, "_work", None) is not None else 0)
        per_gt = max(1, -(-int(n_alleles_total) // max(1, int(n_gt))))
        b = max(1, min(int(want), int(n_gt)))
        while b > 1 and int(self.lib.sonics_genotype_work_bytes_mode(b, b * per_gt, int(max_reps),
                                                                     int(bool(monte_carlo)))) > budget:
            b = (b + 1) // 2
        return b

    def _workspace(self, nbytes):
        w = getattr(self, "_work", None)
        if w is None or w.numel() < nbytes:
            self._work = None
            w = self.torch.empty(int(nbytes), dtype=self.torch.uint8, device=self.device)
            self._work = w
        return w
