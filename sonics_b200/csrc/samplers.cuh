// Exact binomial / Poisson samplers for the PCR kernel.
//
// The reference calls numpy's legacy np.random.binomial (inversion / BTPE) at
// sonics.pyx:354,454,458,462,478 and np.random.poisson (PTRS) at :420.  Those
// are rejection / inversion samplers of the exact distributions; here the same
// distributions are sampled with algorithms that map better onto SIMT lanes:
//   binomial:  n*min(p,1-p) < 10  -> sequential inversion from 0 (BINV)
//              otherwise          -> BTRS transformed rejection (Hoermann, "The generation of
//                                    binomial random variates", J. Stat. Comput. Simul. 46, 1993)
//                                    with the squeeze of BTPE (Kachitvichyanukul & Schmeiser,
//                                    CACM 31, 1988, step 5.2) in front of the Stirling test
//   Poisson:   lam < 10           -> sequential inversion
//              otherwise          -> PTRS (Hoermann 1993)
//
// SIMT notes (ncu, profiles/r01): the cost of a rejection sampler on a warp is set by its
// rarely-taken paths, which run with 1-4 lanes active.  Hence: random words are requested in
// whole Philox blocks at convergent points (rng.cuh); the proposal k = floor(...) is formed in
// split fp32 (integer part of n*p kept exactly, floor by a round-down add) instead of fp64; the
// cheap squeeze decides all but the thin band |ln v' - t| <= rho, so the log1p-conditioned
// Stirling evaluation is needed for well under 1% of trials at large n*p*q; uniforms are built
// with integer ops and reciprocals 1/x of the inversion recurrence come from a shared-memory
// table, which keeps conversions and MUFU ops off the saturated XU pipe.
#pragma once
#include "rng.cuh"

namespace sonics {

#define SONICS_RCP_TABLE 192 // 1/x for x in [0, 192) (inversion recurrences; index 0 unused)

// Stirling tail fc(k) = ln k! - [ (k+1/2) ln(k+1) - (k+1) + 1/2 ln(2 pi) ]
__device__ __forceinline__ float stirling_tail(float k)
{
    const float r = __fdividef(1.0f, k + 1.0f);
    const float r2 = r * r;
    return r * (0.083333333333f - r2 * (0.002777777778f - r2 * 0.000793650794f));
}

// floor(x) for |x| < 2^22 without the XU pipe: round-down add of 1.5 * 2^23
__device__ __forceinline__ float floor_small(float x) { return __fadd_rd(x, 12582912.0f) - 12582912.0f; }

// Binomial(n, p) for 0 < p <= 0.5 and n*p < 10: inversion by sequential search.
__device__ __forceinline__ int binomial_inversion(int n, float p, uint32_t bits, const float *rcp_tab)
{
    const float q = 1.0f - p;
    const float s = __fdividef(p, q);
    const float a = (float)(n + 1) * s;
    float px = __expf((float)n * log1pf(-p)); // q^n
    float u = u01_fine(bits);
    int x = 0;
    const int cap = n < SONICS_RCP_TABLE - 1 ? n : SONICS_RCP_TABLE - 1; // P(X > 160 | np < 10) < 1e-130
    while (u > px && x < cap) {
        ++x;
        u -= px;
        px *= fmaf(a, rcp_tab[x], -s);
    }
    return x;
}

// Binomial(n, p) for 0 < p <= 0.5 and n*p >= 10: BTRS + squeeze.
__device__ __forceinline__ int binomial_btrs(Rng &rng, int n, float p, uint32_t bits_a, uint32_t bits_b)
{
    const float fn = (float)n;
    const float q = 1.0f - p;
    // n*p exactly as hi + lo (n < 2^24, so fma recovers the rounding error of the product)
    const float hi = fn * p;
    const float lo = fmaf(fn, p, -hi);
    const float npq = hi * q;
    const float spq = sqrtf(npq);
    const float b = 1.15f + 2.53f * spq;
    const float rb = __fdividef(1.0f, b);
    const float a = -0.0873f + 0.0248f * b + 0.01f * p;
    const float vr = 0.92f - 4.2f * rb;
    const float alpha = (2.83f + 5.1f * rb) * spq;
    const float ci = floorf(hi);               // integer part of n*p (exact in fp32)
    const float cf = (hi - ci) + lo + 0.5f;    // c = n*p + 0.5 = ci + cf
    const float fm = ci + floorf(cf - 0.5f + p); // m = floor((n+1) p)
    PairSource src(rng, bits_a, bits_b);
    float fk;
    while (true) {
        const float u = u01_open(src.a) - 0.5f;
        const float v = u01_open(src.b);
        const float us = 0.5f - fabsf(u);
        const float rus = __fdividef(1.0f, us);
        // k = floor((2a/us + b) u + c); the fractional work stays below 2^22 in magnitude
        fk = ci + floor_small(fmaf(fmaf(2.0f * a, rus, b), u, cf));
        if (us >= 0.07f && v <= vr)
            break; // quick accept
        if (fk >= 0.0f && fk <= fn) {
            const float A = __logf(v * alpha * __fdividef(1.0f, fmaf(a * rus, rus, b)));
            const float d = fk - fm;
            const float ad = fabsf(d);
            bool decided = false, accept = false;
            if (ad < 0.5f * npq - 1.0f) {
                // BTPE step 5.2: t - rho <= ln f(k)/f(m) <= t + rho
                const float rn = __fdividef(1.0f, npq);
                const float t = -0.5f * d * d * rn;
                const float rho = (ad * rn) * (fmaf(ad, fmaf(ad, 0.33333333f, 0.625f), 0.16666667f) * rn + 0.5f);
                // keep a safety margin for the fp32 evaluation of A, t and rho
                const float eps = 1e-5f * (1.0f + fabsf(t));
                if (A < t - rho - eps) {
                    decided = true;
                    accept = true;
                } else if (A > t + rho + eps) {
                    decided = true;
                }
            }
            if (!decided) {
                // ln f(k)/f(m) with every term conditioned through log1p: each is O(|k-m|)
                const float r = __fdividef(p, q);
                const float nk = fn - fk + 1.0f;
                const float k1 = fk + 1.0f;
                const float t1 = (fm + 0.5f) * log1pf(-d / k1);
                const float t2 = (fn - fm + 0.5f) * log1pf(d / nk);
                const float t3 = d * logf(r * nk / k1);
                const float h = t1 + t2 + t3 + stirling_tail(fm) + stirling_tail(fn - fm) - stirling_tail(fk) -
                                stirling_tail(fn - fk);
                accept = A <= h;
            }
            if (accept)
                break;
        }
        src.advance();
    }
    return (int)fk;
}

// Binomial(n, p), any p in [0, 1].  (bits_a, bits_b): fresh random words for the first trial.
__device__ __forceinline__ int binomial(Rng &rng, int n, float p, uint32_t bits_a, uint32_t bits_b,
                                        const float *rcp_tab)
{
    if (n <= 0 || !(p > 0.0f))
        return 0;
    if (p >= 1.0f)
        return n;
    const bool flip = p > 0.5f;
    const float pp = flip ? 1.0f - p : p;
    int k;
    if ((float)n * pp < 10.0f)
        k = binomial_inversion(n, pp, bits_a, rcp_tab);
    else
        k = binomial_btrs(rng, n, pp, bits_a, bits_b);
    return flip ? n - k : k;
}

// Poisson(lam).  lam arrives as the reference's fp32 `hit` widened to double.  Rare (one draw
// per non-zero bin per reaction), so the acceptance test simply runs in fp64.
__device__ __forceinline__ int poisson(Rng &rng, double lam)
{
    if (!(lam > 0.0))
        return 0;
    const uint4 r0 = rng.block4();
    if (lam < 10.0) {
        float px = __expf(-(float)lam);
        float u = u01_fine(r0.x);
        const float fl = (float)lam;
        int x = 0;
        while (u > px && x < 200) {
            ++x;
            u -= px;
            px *= __fdividef(fl, (float)x);
        }
        return x;
    }
    const double slam = sqrt(lam);
    const double loglam = log(lam);
    const float b = (float)(0.931 + 2.53 * slam);
    const float a = -0.059f + 0.02483f * b;
    const float invalpha = 1.1239f + __fdividef(1.1328f, b - 3.4f);
    const float vr = 0.9277f - __fdividef(3.6224f, b - 2.0f);
    PairSource src(rng, r0.x, r0.y);
    src.sa = r0.z;
    src.sb = r0.w;
    src.spare = 1;
    while (true) {
        const float u = u01_open(src.a) - 0.5f;
        const float v = u01_open(src.b);
        const float us = 0.5f - fabsf(u);
        const double kd = floor((double)(__fdividef(2.0f * a, us) + b) * (double)u + lam + 0.43);
        if (us >= 0.07f && v <= vr)
            return (int)kd;
        if (!(kd < 0.0 || (us < 0.013f && v > us))) {
            const double lhs = (double)(__logf(v) + __logf(invalpha) - __logf(__fdividef(a, us * us) + b));
            const double rhs = -lam + kd * loglam - lgamma(kd + 1.0);
            if (lhs <= rhs)
                return (int)kd;
        }
        src.advance();
    }
}

} // namespace sonics
