"""ctypes front-end of the CPU oracle (oracle/sonics_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py.  The product package
(sonics_b200/) never imports this module.
"""
import ctypes
import os
import subprocess
from concurrent.futures import ThreadPoolExecutor

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libsonics_oracle.so")
MAX_ALLELE = 500


class OracleCfg(ctypes.Structure):
    _fields_ = [
        ("n_cycles", ctypes.c_int32),
        ("capture_cycle", ctypes.c_int32),
        ("floor_opt", ctypes.c_int32),
        ("random_start", ctypes.c_int32),
        ("up_preference", ctypes.c_int32),
        ("rng_mode", ctypes.c_int32),
        ("force_second", ctypes.c_int32),
        ("with_readout", ctypes.c_int32),
        ("start_copies", ctypes.c_int64),
        ("ranges", ctypes.c_double * 8),
        ("noise_coef", ctypes.c_float),
        ("_pad", ctypes.c_int32),
    ]


RECORD_DTYPE = np.dtype(
    [
        ("identity", "<f8"),
        ("r_squared", "<f8"),
        ("lnL", "<f8"),
        ("first", "<i4"),
        ("second", "<i4"),
        ("down", "<f8"),
        ("up", "<f8"),
        ("capture", "<f8"),
        ("efficiency", "<f8"),
        ("steps_stutter", "<i8"),
        ("steps_plain", "<i8"),
        ("n_poisson", "<i8"),
        ("n_binom_btpe", "<i8"),
        ("n_binom_inv", "<i8"),
        ("n_binom_zero", "<i8"),
        ("live_lo", "<i4"),
        ("live_hi", "<i4"),
    ],
    align=True,
)


def build(force=False):
    """Compile oracle/libsonics_oracle.so with gcc (no-op when up to date)."""
    src = os.path.join(HERE, "sonics_oracle.c")
    if force or not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", HERE, "-s"])
    return LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(LIB_PATH)
        L.oracle_gammln.restype = ctypes.c_double
        L.oracle_gammln.argtypes = [ctypes.c_double]
        L.oracle_factln.restype = ctypes.c_double
        L.oracle_factln.argtypes = [ctypes.c_int32]
        L.oracle_mvhyperg.restype = ctypes.c_double
        L.oracle_mvhyperg.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
        L.oracle_rng_new.restype = ctypes.c_void_p
        L.oracle_rng_new.argtypes = [ctypes.c_uint32]
        L.oracle_rng_free.argtypes = [ctypes.c_void_p]
        L.oracle_rng_seed.argtypes = [ctypes.c_void_p, ctypes.c_uint32]
        L.oracle_rng_uint32.restype = ctypes.c_uint32
        L.oracle_rng_uint32.argtypes = [ctypes.c_void_p]
        L.oracle_rng_double.restype = ctypes.c_double
        L.oracle_rng_double.argtypes = [ctypes.c_void_p]
        L.oracle_uniform.restype = ctypes.c_double
        L.oracle_uniform.argtypes = [ctypes.c_void_p, ctypes.c_double, ctypes.c_double]
        L.oracle_randint.restype = ctypes.c_int64
        L.oracle_randint.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64]
        L.oracle_binomial.restype = ctypes.c_int64
        L.oracle_binomial.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_double]
        L.oracle_poisson.restype = ctypes.c_int64
        L.oracle_poisson.argtypes = [ctypes.c_void_p, ctypes.c_double]
        L.oracle_run.restype = ctypes.c_int
        L.oracle_run.argtypes = [ctypes.POINTER(OracleCfg), ctypes.c_void_p, ctypes.c_int, ctypes.c_uint32,
                                 ctypes.c_int64, ctypes.c_void_p]
        L.oracle_replay_script.restype = ctypes.c_int
        L.oracle_replay_script.argtypes = [ctypes.POINTER(OracleCfg), ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p,
                                           ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_int64,
                                           ctypes.c_void_p, ctypes.c_void_p, ctypes.POINTER(ctypes.c_int64),
                                           ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_double),
                                           ctypes.POINTER(ctypes.c_int64)]
        L.oracle_record_script.restype = ctypes.c_int
        L.oracle_record_script.argtypes = [ctypes.POINTER(OracleCfg), ctypes.c_void_p, ctypes.c_int, ctypes.c_uint32,
                                           ctypes.c_void_p, ctypes.c_int64, ctypes.POINTER(ctypes.c_int64),
                                           ctypes.c_void_p, ctypes.c_void_p]
        L.oracle_one_repeat.restype = ctypes.c_int
        L.oracle_one_repeat.argtypes = [ctypes.c_void_p, ctypes.POINTER(OracleCfg), ctypes.c_void_p, ctypes.c_int,
                                        ctypes.c_void_p, ctypes.c_void_p]
        L.oracle_replay_script_where.restype = ctypes.c_int
        L.oracle_replay_script_where.argtypes = [ctypes.POINTER(OracleCfg), ctypes.c_void_p, ctypes.c_int,
                                                 ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p,
                                                 ctypes.c_int64, ctypes.POINTER(ctypes.c_double),
                                                 ctypes.POINTER(ctypes.c_int64), ctypes.POINTER(ctypes.c_double)]
        L.oracle_descriptors.restype = None
        L.oracle_descriptors.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int,
                                         ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_double)]
        assert L.oracle_sizeof_cfg() == ctypes.sizeof(OracleCfg), (L.oracle_sizeof_cfg(), ctypes.sizeof(OracleCfg))
        assert L.oracle_sizeof_record() == RECORD_DTYPE.itemsize, (L.oracle_sizeof_record(), RECORD_DTYPE.itemsize)
        _lib = L
    return _lib


def parse_reads(genot_input):
    """Oracle-side restatement of get_alleles (sonics.pyx:17-31)."""
    alleles = np.zeros(MAX_ALLELE, dtype=np.int64)
    for f in genot_input.split(";"):
        if f != "":
            a, n = [int(x) for x in f.split("|")]
            if a > 0:
                alleles[a] = n
    return alleles


def auto_noise_coef(alleles, noise_threshold=0.05):
    """sonics:407-416."""
    nz = alleles[alleles.nonzero()]
    frac = nz.min() / alleles.sum() * np.count_nonzero(alleles)
    noise_frac = frac / (frac + 1)
    return float(nz.min() / alleles.sum()) if noise_frac < noise_threshold else 0.0


def make_cfg(n_cycles=24, capture_cycle=13, floor_opt=5, random_start=False, up_preference=False,
             start_copies=300000, down=(0.0, 0.1), up=(0.0, 0.1), capture=(-0.25, 0.25),
             efficiency=(0.001, 0.1), noise_coef=0.0, rng_mode=0, force_second=-1, with_readout=True):
    cfg = OracleCfg()
    cfg.n_cycles = n_cycles
    cfg.capture_cycle = capture_cycle
    cfg.floor_opt = floor_opt
    cfg.random_start = int(bool(random_start))
    cfg.up_preference = int(bool(up_preference))
    cfg.rng_mode = rng_mode
    cfg.force_second = force_second
    cfg.with_readout = int(bool(with_readout))
    cfg.start_copies = start_copies
    for i, v in enumerate([down[0], down[1], up[0], up[1], capture[0], capture[1], efficiency[0], efficiency[1]]):
        cfg.ranges[i] = float(v)
    cfg.noise_coef = float(noise_coef)
    return cfg


def run(cfg, alleles, seed, n):
    """n consecutive one_repeat() calls on one MT19937 stream seeded like np.random.seed(seed)."""
    alleles = np.ascontiguousarray(alleles, dtype=np.int64)
    out = np.zeros(n, dtype=RECORD_DTYPE)
    rc = lib().oracle_run(ctypes.byref(cfg), alleles.ctypes.data, int(alleles.shape[0]), int(seed) & 0xFFFFFFFF,
                          int(n), out.ctypes.data)
    if rc == -1:
        raise Exception("Less then two alleles as starting conditions! Aborting.")
    return out


SCRIPT_DTYPE = np.dtype([("kind", "<i4"), ("n", "<i4"), ("p", "<f4"), ("x", "<i4")])  # oracle_script_entry


def replay_script(cfg, alleles, params4, start_idx, script):
    """One one_repeat() (no readout) of the reference's loop on recorded draw outcomes.

    params4 = (down, up, capture, efficiency); start_idx = the randint() results of the start-allele choice
    (indices into the non-zero input alleles); script = SCRIPT_DTYPE records in request order.  Returns
    (rc, record, pool[int64, max_allele], used, max_rel_dp, max_abs_dlam, err_pos); rc 0 = the script was
    consistent with every request the loop made (kind and n equal, outcomes in range)."""
    alleles = np.ascontiguousarray(alleles, dtype=np.int64)
    params4 = np.ascontiguousarray(params4, dtype=np.float64)
    start_idx = np.ascontiguousarray(start_idx, dtype=np.int64)
    script = np.ascontiguousarray(script, dtype=SCRIPT_DTYPE)
    rec = np.zeros(1, dtype=RECORD_DTYPE)
    pool = np.zeros(alleles.shape[0], dtype=np.int64)
    used, err_pos = ctypes.c_int64(0), ctypes.c_int64(0)
    dp, dlam = ctypes.c_double(0), ctypes.c_double(0)
    rc = lib().oracle_replay_script(ctypes.byref(cfg), alleles.ctypes.data, int(alleles.shape[0]),
                                    params4.ctypes.data, start_idx.ctypes.data, int(start_idx.shape[0]),
                                    script.ctypes.data, int(script.shape[0]), rec.ctypes.data, pool.ctypes.data,
                                    ctypes.byref(used), ctypes.byref(dp), ctypes.byref(dlam), ctypes.byref(err_pos))
    return rc, rec[0], pool, used.value, dp.value, dlam.value, err_pos.value


def replay_script_where(cfg, alleles, params4, start_idx, script):
    """(max relative deviation of a binomial probability, index of that request, the reference's p there)."""
    alleles = np.ascontiguousarray(alleles, dtype=np.int64)
    params4 = np.ascontiguousarray(params4, dtype=np.float64)
    start_idx = np.ascontiguousarray(start_idx, dtype=np.int64)
    script = np.ascontiguousarray(script, dtype=SCRIPT_DTYPE)
    dp, pos, pref = ctypes.c_double(0), ctypes.c_int64(0), ctypes.c_double(0)
    lib().oracle_replay_script_where(ctypes.byref(cfg), alleles.ctypes.data, int(alleles.shape[0]),
                                     params4.ctypes.data, start_idx.ctypes.data, int(start_idx.shape[0]),
                                     script.ctypes.data, int(script.shape[0]), ctypes.byref(dp), ctypes.byref(pos),
                                     ctypes.byref(pref))
    return dp.value, pos.value, pref.value


def record_script(cfg, alleles, seed, cap=1 << 16):
    """One sampled one_repeat() (no readout) -> (record, pool, script of its draws in request order)."""
    alleles = np.ascontiguousarray(alleles, dtype=np.int64)
    script = np.zeros(cap, dtype=SCRIPT_DTYPE)
    rec = np.zeros(1, dtype=RECORD_DTYPE)
    pool = np.zeros(alleles.shape[0], dtype=np.int64)
    n = ctypes.c_int64(0)
    rc = lib().oracle_record_script(ctypes.byref(cfg), alleles.ctypes.data, int(alleles.shape[0]),
                                    int(seed) & 0xFFFFFFFF, script.ctypes.data, cap, ctypes.byref(n), rec.ctypes.data,
                                    pool.ctypes.data)
    if rc == -1:
        raise Exception("Less then two alleles as starting conditions! Aborting.")
    assert n.value <= cap, "script capacity %d < %d draws" % (cap, n.value)
    return rec[0], pool, script[:n.value]


def run_parallel(cfg, alleles, seed, n, threads=None, chunk=2000):
    """Same distribution as run(), split over host threads (independent seeds per chunk)."""
    threads = threads or os.cpu_count() or 1
    jobs = []
    i = 0
    k = 0
    while i < n:
        m = min(chunk, n - i)
        jobs.append((k, m))
        i += m
        k += 1
    ss = np.random.SeedSequence(int(seed))
    seeds = [int(s.generate_state(1)[0]) for s in ss.spawn(len(jobs))]
    with ThreadPoolExecutor(threads) as ex:
        parts = list(ex.map(lambda j: run(cfg, alleles, seeds[j[0]], j[1]), jobs))
    return np.concatenate(parts) if parts else np.zeros(0, dtype=RECORD_DTYPE)


def mvhyperg(x, color):
    x = np.ascontiguousarray(x, dtype=np.int64)
    color = np.ascontiguousarray(color, dtype=np.int64)
    return float(lib().oracle_mvhyperg(x.ctypes.data, color.ctypes.data, int(x.shape[0])))


def descriptors(alleles, readout):
    """(identity, r_squared) of a sequencing readout against the input reads (sonics.pyx:373-384, rsq :273-291)."""
    alleles = np.ascontiguousarray(alleles, dtype=np.int64)
    readout = np.ascontiguousarray(readout, dtype=np.int64)
    assert alleles.shape == readout.shape
    ident, r2 = ctypes.c_double(0), ctypes.c_double(0)
    lib().oracle_descriptors(alleles.ctypes.data, readout.ctypes.data, int(alleles.shape[0]), ctypes.byref(ident),
                             ctypes.byref(r2))
    return ident.value, r2.value


def label(rec):
    a, b = int(rec["first"]), int(rec["second"])
    return "%d/%d" % (min(a, b), max(a, b))
