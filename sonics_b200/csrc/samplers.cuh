// Exact binomial / Poisson samplers for the PCR kernel.
//
// The reference calls numpy's legacy np.random.binomial (inversion / BTPE) at
// sonics.pyx:354,454,458,462,478 and np.random.poisson (PTRS) at :420.  Those
// are rejection / inversion samplers of the exact distributions; here the same
// distributions are sampled with algorithms that map better onto SIMT lanes:
//   binomial:  n*min(p,1-p) < 10  -> sequential inversion from 0 (BINV)
//              otherwise          -> BTRS transformed rejection with squeeze
//                                    (Hoermann, "The generation of binomial random
//                                    variates", J. Stat. Comput. Simul. 46, 1993)
//   Poisson:   lam < 10           -> sequential inversion
//              otherwise          -> PTRS (Hoermann, "The transformed rejection method
//                                    for generating Poisson random variables", 1993)
// Arithmetic: proposals k = floor(...) are formed in fp64 (counts reach ~1e7, past
// fp32's integer range of exact .5 steps); acceptance tests use log1p-conditioned
// fp32 (binomial: every term is O(|k-m|), relative error 1e-7 => acceptance error
// <= ~2e-4 in the far tail) or fp64 (Poisson, whose terms are O(lam log lam)).
#pragma once
#include "rng.cuh"

namespace sonics {

// Stirling tail fc(k) = ln k! - [ (k+1/2) ln(k+1) - (k+1) + 1/2 ln(2 pi) ]
__device__ __forceinline__ float stirling_tail(float k)
{
    const float r = __frcp_rn(k + 1.0f);
    const float r2 = r * r;
    return r * (0.083333333333f - r2 * (0.002777777778f - r2 * 0.000793650794f));
}

// Binomial(n, p) for 0 < p <= 0.5 and n*p < 10: inversion by sequential search.
__device__ __forceinline__ int binomial_inversion(Rng &rng, int n, float p)
{
    const float q = 1.0f - p;
    const float s = __fdividef(p, q);
    const float a = (float)(n + 1) * s;
    float px = __expf((float)n * log1pf(-p)); // q^n
    float u = rng.uniform_fine();
    int x = 0;
    const int cap = n < 160 ? n : 160; // P(X > 160 | np < 10) < 1e-130
    while (u > px && x < cap) {
        ++x;
        u -= px;
        px *= (__fdividef(a, (float)x) - s);
    }
    return x;
}

// Binomial(n, p) for 0 < p <= 0.5 and n*p >= 10: BTRS.
__device__ __forceinline__ int binomial_btrs(Rng &rng, int n, float p)
{
    const float fn = (float)n;
    const float q = 1.0f - p;
    const float spq = sqrtf(fn * p * q);
    const float b = 1.15f + 2.53f * spq;
    const float a = -0.0873f + 0.0248f * b + 0.01f * p;
    const double c = (double)n * (double)p + 0.5;
    const float vr = 0.92f - __fdividef(4.2f, b);
    const float alpha = (2.83f + __fdividef(5.1f, b)) * spq;
    const float r = __fdividef(p, q);
    const float fm = floorf((fn + 1.0f) * p);
    int k;
    while (true) {
        const float u = rng.uniform_open() - 0.5f;
        float v = rng.uniform_open();
        const float us = 0.5f - fabsf(u);
        const double kd = floor((double)(__fdividef(2.0f * a, us) + b) * (double)u + c);
        if (us >= 0.07f && v <= vr) {
            k = (int)kd;
            break;
        }
        if (kd < 0.0 || kd > (double)n)
            continue;
        k = (int)kd;
        const float fk = (float)k;
        v = __logf(v * __fdividef(alpha, __fdividef(a, us * us) + b));
        const float d = fk - fm;
        const float nk = fn - fk + 1.0f;
        const float k1 = fk + 1.0f;
        // ln f(k)/f(m), each term conditioned through log1p (see header comment)
        const float t1 = (fm + 0.5f) * log1pf(-d / k1);
        const float t2 = (fn - fm + 0.5f) * log1pf(d / nk);
        const float t3 = d * logf(r * nk / k1);
        const float h = t1 + t2 + t3 + stirling_tail(fm) + stirling_tail(fn - fm) - stirling_tail(fk) -
                        stirling_tail(fn - fk);
        if (v <= h)
            break;
    }
    return k;
}

// Binomial(n, p), any p in [0, 1].
__device__ __forceinline__ int binomial(Rng &rng, int n, float p)
{
    if (n <= 0 || !(p > 0.0f))
        return 0;
    if (p >= 1.0f)
        return n;
    const bool flip = p > 0.5f;
    const float pp = flip ? 1.0f - p : p;
    int k;
    if ((float)n * pp < 10.0f)
        k = binomial_inversion(rng, n, pp);
    else
        k = binomial_btrs(rng, n, pp);
    return flip ? n - k : k;
}

// Poisson(lam).  lam arrives as the reference's fp32 `hit` widened to double.
__device__ __forceinline__ int poisson(Rng &rng, double lam)
{
    if (!(lam > 0.0))
        return 0;
    if (lam < 10.0) {
        float px = __expf(-(float)lam);
        float u = rng.uniform_fine();
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
    while (true) {
        const float u = rng.uniform_open() - 0.5f;
        const float v = rng.uniform_open();
        const float us = 0.5f - fabsf(u);
        const double kd = floor((double)(__fdividef(2.0f * a, us) + b) * (double)u + lam + 0.43);
        if (us >= 0.07f && v <= vr)
            return (int)kd;
        if (kd < 0.0 || (us < 0.013f && v > us))
            continue;
        const double lhs = (double)(__logf(v) + __logf(invalpha) - __logf(__fdividef(a, us * us) + b));
        const double rhs = -lam + kd * loglam - lgamma(kd + 1.0);
        if (lhs <= rhs)
            return (int)kd;
    }
}

} // namespace sonics
