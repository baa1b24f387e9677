"""ctypes binding of include/sonics_b200.h.

The product path has NO fallback: if libsonics_b200.so is missing or no sm_100
device is present, loading / sonics_create raises.
"""
import ctypes
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libsonics_b200.so")

SENTINEL = -999999.0
KERNEL_AUTO, KERNEL_PERSISTENT, KERNEL_LOCKSTEP = 0, 1, 2
FILTER_PASS, FILTER_MWU_TEST, FILTER_BEST_RATIO, FILTER_PERCENTILE_RATIO, FILTER_NO_SUCCESS = 1, 2, 4, 8, 16


class SonicsError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("sonics_b200 error %d: %s" % (code, msg))
        self.code = code


class SonicsParams(ctypes.Structure):
    _fields_ = [
        ("n_cycles", ctypes.c_int32),
        ("capture_cycle", ctypes.c_int32),
        ("floor_opt", ctypes.c_int32),
        ("random_start", ctypes.c_int32),
        ("up_preference", ctypes.c_int32),
        ("start_copies", ctypes.c_int32),
        ("min_sim", ctypes.c_int32),
        ("monte_carlo", ctypes.c_int32),
        ("down", ctypes.c_double * 2),
        ("up", ctypes.c_double * 2),
        ("capture", ctypes.c_double * 2),
        ("efficiency", ctypes.c_double * 2),
        ("padjust", ctypes.c_double),
        ("lnL_threshold", ctypes.c_double),
        ("n_add_ratios", ctypes.c_int32),
        ("_pad", ctypes.c_int32),
        ("add_ratios", ctypes.c_double * 8),
    ]


class SonicsVerdict(ctypes.Structure):
    _fields_ = [
        ("allele_lo", ctypes.c_int32),
        ("allele_hi", ctypes.c_int32),
        ("filter", ctypes.c_int32),
        ("run_reps", ctypes.c_int32),
        ("min_sim", ctypes.c_int32),
        ("best_sim", ctypes.c_int32),
        ("n_groups", ctypes.c_int32),
        ("p_clamped", ctypes.c_int32),
        ("lnL", ctypes.c_double),
        ("identity", ctypes.c_double),
        ("r_squared", ctypes.c_double),
        ("high_pval", ctypes.c_double),
        ("best_lnL_ratio", ctypes.c_double),
        ("percentile_lnL_ratio", ctypes.c_double),
        ("identity_q75", ctypes.c_double),
        ("r_squared_q75", ctypes.c_double),
        ("lnL_q75", ctypes.c_double),
        ("add_ratio", ctypes.c_double * 8),
    ]


# every symbol include/sonics_b200.h declares: name -> (restype, argtypes)
_P = ctypes.c_void_p
_I = ctypes.c_int
_I64 = ctypes.c_int64
_U64 = ctypes.c_uint64
_SZ = ctypes.c_size_t
PROTOTYPES = {
    "sonics_abi_version": (_I, []),
    "sonics_last_error": (ctypes.c_char_p, []),
    "sonics_create": (_I, [_I, ctypes.POINTER(_P)]),
    "sonics_destroy": (None, [_P]),
    "sonics_sync": (_I, [_P, _P]),
    "sonics_mvhyperg": (_I, [_P, _P, _P, _I, _I, _P, _P, _SZ]),
    "sonics_table_bytes": (_SZ, [_I, _I]),
    "sonics_table_build": (_I, [_P, ctypes.POINTER(SonicsParams), _I, _P, _P, _P, _P, _P, _P, _SZ, _P]),
    "sonics_table_max_window": (_I, [_P]),
    "sonics_table_release": (_I, [_P, _P]),
    "sonics_simulate": (_I, [_P, ctypes.POINTER(SonicsParams), _P, _I, _P, _I, _I64, _I64, _U64, _I64, _I64, _P, _P,
                             _P, _P, _P, _P, _SZ, _P]),
    "sonics_scratch_bytes": (_SZ, [_I64, _I]),
    "sonics_scratch_bytes_kernel": (_SZ, [_I64, _I, _I]),
    "sonics_set_kernel": (_I, [_P, _I]),
    "sonics_set_sort_bits": (_I, [_P, _I, _I, _I]),
    "sonics_set_sort_mode": (_I, [_P, _I]),
    "sonics_set_samplers": (_I, [_P, _I, ctypes.c_double]),
    "sonics_debug_readout": (_I, [_P, _P]),
    "sonics_evaluate": (_I, [_P, ctypes.POINTER(SonicsParams), _P, _I, _P, _I, _I64, _I64, _P, _P, _P, _P, _I, _P, _P,
                             _SZ, _P]),
    "sonics_stats_scratch_bytes": (_SZ, [_I64, _I, _I]),
    "sonics_stats_scratch_bytes_batch": (_SZ, [_I, _I64, _I, _I]),
    "sonics_replay_best": (_I, [_P, ctypes.POINTER(SonicsParams), _P, _I, _P, _I, _U64, _P, _P, _SZ, _P]),
    "sonics_genotype_work_bytes": (_SZ, [_I, _I, _I]),
    "sonics_genotype_work_bytes_mode": (_SZ, [_I, _I, _I, _I]),
    "sonics_genotype_work_bytes_for": (_SZ, [ctypes.POINTER(SonicsParams), _I, _P, _I]),
    "sonics_genotype_batch": (_I, [_P, ctypes.POINTER(SonicsParams), _I, _P, _P, _P, _P, _P, _I, _U64, _P, _P, _SZ,
                                   _P]),
    "sonics_debug_counters": (_I, [_P, _P]),
    "sonics_debug_trace": (_I, [_P, _P, ctypes.c_int32, _P]),
    "sonics_debug_philox": (_I, [_P, _U64, _U64, ctypes.c_uint32, ctypes.c_uint32, _I, _P]),
    "sonics_debug_pairs": (_I, [_P, _U64, _U64, ctypes.c_uint32, _I, _P]),
    "sonics_debug_sample": (_I, [_P, _I, _I, ctypes.c_double, _I64, _U64, _P]),
    "sonics_launch_count": (_I64, [_P]),
    "sonics_vcf_parse": (_I, [ctypes.c_char_p, _I, _I, ctypes.c_double, _I, _U64, ctypes.POINTER(_P)]),
    "sonics_vcf_parse_range": (_I, [ctypes.c_char_p, _I64, _I64, _I, _I, ctypes.c_double, _I, _U64,
                                    ctypes.POINTER(_P)]),
    "sonics_vcf_free": (None, [_P]),
    "sonics_vcf_count": (_I64, [_P, _I]),
    "sonics_vcf_fill": (_I, [_P, _P, _P, _P, _P, _P]),
    "sonics_vcf_uids": (_I, [_P, _I64, _P]),
    "sonics_vcf_lines_before": (_I, [ctypes.c_char_p, _I64, ctypes.POINTER(_I64)]),
    "sonics_vcf_count_lines": (_I, [ctypes.c_char_p, _I64, _I64, ctypes.POINTER(_I64)]),
    "sonics_vcf_warning": (ctypes.c_char_p, [_P, _I64]),
    "sonics_rows_write": (_I, [ctypes.c_char_p, _P, _P, _I64, _I, ctypes.c_char_p, ctypes.c_char_p, ctypes.c_char_p]),
    "sonics_rows_write_part": (_I, [ctypes.c_char_p, _P, _P, _I64, _I, ctypes.c_char_p, ctypes.c_char_p,
                                    ctypes.c_char_p, _I]),
    "sonics_rows_concat": (_I, [ctypes.c_char_p, ctypes.POINTER(ctypes.c_char_p), _I]),
}

_lib = None


def load():
    """Load libsonics_b200.so (raises ImportError when it has not been built)."""
    global _lib
    if _lib is None:
        path = os.environ.get("SONICS_B200_LIB", LIB_PATH)  # another build of the same ABI (kernel A/B runs)
        if not os.path.exists(path):
            raise ImportError(
                "sonics_b200: %s is missing -- build it with `python -m sonics_b200.build` "
                "(there is no CPU fallback)" % path)
        lib = ctypes.CDLL(path)
        for name, (res, args) in PROTOTYPES.items():
            fn = getattr(lib, name)  # AttributeError if the symbol is not exported
            fn.restype = res
            fn.argtypes = args
        if lib.sonics_abi_version() != 1:
            raise ImportError("sonics_b200: ABI version mismatch")
        _lib = lib
    return _lib


def check(rc):
    if rc != 0:
        raise SonicsError(rc, load().sonics_last_error().decode("utf-8", "replace"))
