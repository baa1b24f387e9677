// Multivariate-hypergeometric log-likelihood in fp64, restating the reference's
// Fortran (pymc_extracted.f) operation for operation so that the deterministic
// lnL of a fixed PCR pool agrees with it to ~1e-10 relative:
//   gammln   pymc_extracted.f:6-31   Numerical-Recipes 6-term Lanczos
//   factln   pymc_extracted.f:34-61  ln n! = gammln(n+1); n<0 -> -DBL_MAX
//   mvhyperg pymc_extracted.f:63-109 sum_i ln C(color_i, x_i) - ln C(total, d)
// Every fp64 operation uses the round-to-nearest intrinsics so nvcc cannot
// contract a*b+c into an FMA: the Fortran build evaluates plain IEEE mul/add,
// and terms of size 1e7 cancelling to O(10) make the contraction visible.
// The one thing that can differ from the reference is the last bit of log().
#pragma once
#include <math.h>

namespace sonics {

#define SONICS_NEG_INFINITY (-1.7976931348623157e308)

__device__ __forceinline__ double gammln_dev(double xx)
{
    const double x = xx;
    double y = x;
    double tmp = __dadd_rn(x, 5.5);
    tmp = __dsub_rn(__dmul_rn(__dadd_rn(x, 0.5), log(tmp)), tmp);
    double ser = 1.000000000190015;
    y = __dadd_rn(y, 1.0);
    ser = __dadd_rn(ser, __ddiv_rn(76.18009172947146, y));
    y = __dadd_rn(y, 1.0);
    ser = __dadd_rn(ser, __ddiv_rn(-86.50532032941677, y));
    y = __dadd_rn(y, 1.0);
    ser = __dadd_rn(ser, __ddiv_rn(24.01409824083091, y));
    y = __dadd_rn(y, 1.0);
    ser = __dadd_rn(ser, __ddiv_rn(-1.231739572450155, y));
    y = __dadd_rn(y, 1.0);
    ser = __dadd_rn(ser, __ddiv_rn(.1208650973866179e-2, y));
    y = __dadd_rn(y, 1.0);
    ser = __dadd_rn(ser, __ddiv_rn(-.5395239384953e-5, y));
    return __dadd_rn(tmp, log(__ddiv_rn(__dmul_rn(2.5066282746310005, ser), x)));
}

// ln n!.  The reference's gammln(1.0) evaluates to exactly 0.0 in IEEE fp64
// (probed); n == 0 is pinned to that value instead of relying on
// the last bit of the device log().
__device__ __forceinline__ double factln_dev(int n)
{
    if (n < 0)
        return SONICS_NEG_INFINITY;
    if (n == 0)
        return 0.0;
    return gammln_dev(__dadd_rn((double)n, 1.0));
}

} // namespace sonics
