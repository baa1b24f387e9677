"""sonics_b200 -- B200-native (sm_100a) Monte-Carlo engine of SONiCS.

Host-side mirror of the reference's module interface lives in sonics_b200.sonics
(get_alleles, monte_carlo, one_repeat, ...); the CUDA kernels and the C-ABI are in
sonics_b200/csrc and include/sonics_b200.h.  There is no CPU fallback.
"""
__version__ = "0.1.0"
