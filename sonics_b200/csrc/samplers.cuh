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
//
// Rounding.  The same draw is made by two differently structured kernels (k_pcr, k_pcr_ls) that must agree bit
// for bit, so wherever a product feeds a sum the fused / unfused choice is written out (fmaf, __fmul_rn,
// __fadd_rn): left to the compiler's contraction it came out differently in the two kernels for one PTRS
// expression (3.5e-5 of the simulations ended with another pool; r02 diagnosis).
#pragma once
#include "rng.cuh"

namespace sonics {

#define SONICS_RCP_TABLE 192 // 1/x for x in [0, 192) (inversion recurrences; index 0 unused)

// Single-instruction MUFU approximations.  __fdividef / __logf / sqrtf wrap the same units in range scaling
// for denormal and huge operands (5-8 instructions each; the BTRS quick path is ~17 instructions long and
// contains one reciprocal).  Every operand here is a probability, a count below 2^31 or a uniform
// >= 2^-24, far from both ends of the range; the accuracy class is the same (<= 2 ulp).
__device__ __forceinline__ float rcp_fast(float x)
{
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float log_fast(float x)
{
    float r;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r * 0.69314718056f;
}
__device__ __forceinline__ float sqrt_fast(float x)
{
    float r;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

// Stirling tail fc(k) = ln k! - [ (k+1/2) ln(k+1) - (k+1) + 1/2 ln(2 pi) ]
__device__ __forceinline__ float stirling_tail(float k)
{
    const float r = rcp_fast(k + 1.0f);
    const float r2 = r * r;
    return r * fmaf(-r2, fmaf(-r2, 0.000793650794f, 0.002777777778f), 0.083333333333f);
}

// floor(x) for |x| < 2^22 without the XU pipe: round-down add of 1.5 * 2^23
__device__ __forceinline__ float floor_small(float x) { return __fadd_rd(x, 12582912.0f) - 12582912.0f; }

// ---- building blocks -------------------------------------------------------------------------
// The samplers are split into setup / one-trial / deferred-test pieces so that K1's per-lane state
// machine can interleave them across lanes (one rejection trial or one chunk of inversion steps per
// pass, the rare Stirling test batched over several lanes).  binomial() / poisson() below compose
// the same pieces in lockstep form and are what the exact-pmf tests exercise.
//
// A lane has at most one draw in flight, so the three samplers share one register block:
//   BTRS      fn p npq a b vr alpha ci cf fm | fk = candidate, A = ln of the scaled uniform
//   inversion a = (n+1) p/q, b = p/q         | fk = u (remaining uniform), A = px, ix = x, icap = cap
//   PTRS      fn = lam (the reference's fp32 `hit`), a b vr, alpha = 1/alpha | ix = candidate
struct DrawState {
    float fn, p, npq, a, b, vr, alpha, ci, cf, fm, fk, A;
    int ix, icap;
};

// ---- inversion: Binomial(n, p), p <= 0.5, n*p < 10 ----
__device__ __forceinline__ void inv_setup(DrawState &v, int n, float p)
{
    const float q = 1.0f - p;
    v.b = p * rcp_fast(q);
    v.a = (float)(n + 1) * v.b;
    v.A = __expf((float)n * log1pf(-p)); // px = q^n
    v.icap = n < SONICS_RCP_TABLE - 1 ? n : SONICS_RCP_TABLE - 1; // P(X > 160 | np < 10) < 1e-130
    v.ix = 0;
    v.fk = 0.0f;
}
__device__ __forceinline__ void inv_start(DrawState &v, uint32_t bits) { v.fk = u01_fine(bits); }
// the sequential search, to its end (at most icap < SONICS_RCP_TABLE steps; n*p < 10 makes ~2 typical):
// result v.ix.  No step counter and a running table pointer: the loop runs with ~3 lanes, so every
// instruction in it is paid ten times over.
__device__ __forceinline__ void inv_run(DrawState &v, const float *rcp_tab)
{
    // 32-bit shared-memory addresses: a generic float* costs two extra 64-bit adds per step
    const uint32_t base = (uint32_t)__cvta_generic_to_shared(rcp_tab);
    uint32_t r = base + 4u * (uint32_t)v.ix;
    const uint32_t rend = base + 4u * (uint32_t)v.icap;
    float u = v.fk, px = v.A;
    while (u > px && r < rend) {
        r += 4u;
        u -= px;
        float rc;
        asm volatile("ld.shared.f32 %0, [%1];" : "=f"(rc) : "r"(r));
        px *= fmaf(v.a, rc, -v.b);
    }
    v.ix = (int)((r - base) >> 2);
}
// ---- BTRS: Binomial(n, p), p <= 0.5, n*p >= 10 ----
__device__ __forceinline__ void btrs_setup(DrawState &s, int n, float p)
{
    s.fn = (float)n;
    s.p = p;
    const float q = 1.0f - p;
    // n*p exactly as hi + lo (n < 2^24, so fma recovers the rounding error of the product)
    const float hi = s.fn * p;
    const float lo = fmaf(s.fn, p, -hi);
    s.npq = hi * q;
    const float spq = sqrt_fast(s.npq);
    s.b = fmaf(2.53f, spq, 1.15f);
    const float rb = rcp_fast(s.b);
    s.a = fmaf(0.01f, p, fmaf(0.0248f, s.b, -0.0873f));
    s.vr = fmaf(-4.2f, rb, 0.92f);
    s.alpha = fmaf(5.1f, rb, 2.83f) * spq;
    s.ci = floorf(hi);                        // integer part of n*p (exact in fp32)
    s.cf = (hi - s.ci) + lo + 0.5f;           // c = n*p + 0.5 = ci + cf
    s.fm = s.ci + floorf(s.cf - 0.5f + p);    // m = floor((n+1) p)
}

// One rejection trial.  0 = rejected, 1 = accepted (result s.fk), 2 = undecided: s.fk / s.A hold the
// candidate for btrs_full().
__device__ __forceinline__ int btrs_trial(DrawState &s, uint32_t bits_a, uint32_t bits_b)
{
    const float u = u01_open(bits_a) - 0.5f;
    const float v = u01_open(bits_b);
    const float us = 0.5f - fabsf(u);
    const float rus = rcp_fast(us);
    // k = floor((2a/us + b) u + c); the fractional work stays below 2^22 in magnitude
    s.fk = s.ci + floor_small(fmaf(fmaf(2.0f * s.a, rus, s.b), u, s.cf));
    if (us >= 0.07f && v <= s.vr)
        return 1; // quick accept
    if (!(s.fk >= 0.0f && s.fk <= s.fn))
        return 0;
    s.A = log_fast(v * s.alpha * rcp_fast(fmaf(s.a * rus, rus, s.b)));
    const float d = s.fk - s.fm;
    const float ad = fabsf(d);
    if (ad < fmaf(0.5f, s.npq, -1.0f)) {
        // BTPE step 5.2: t - rho <= ln f(k)/f(m) <= t + rho
        const float rn = rcp_fast(s.npq);
        const float t = __fmul_rn(-0.5f * d * d, rn);
        const float rho = __fmul_rn(ad * rn, fmaf(fmaf(ad, fmaf(ad, 0.33333333f, 0.625f), 0.16666667f), rn, 0.5f));
        const float eps = fmaf(1e-5f, fabsf(t), 1e-5f); // margin for the fp32 evaluation of A, t and rho
        if (s.A < t - rho - eps)
            return 1;
        if (s.A > t + rho + eps)
            return 0;
    }
    return 2;
}

// The exact test for an undecided candidate: A <= ln f(k)/f(m), every term conditioned through
// log1p so that each is O(|k-m|) (relative error 1e-7 per term => acceptance error <= ~2e-4 in the
// far tail, on the ~2 % of trials that get here).
__device__ __forceinline__ bool btrs_full(const DrawState &s)
{
    const float q = 1.0f - s.p;
    const float r = s.p * rcp_fast(q);
    const float d = s.fk - s.fm;
    const float nk = s.fn - s.fk + 1.0f;
    const float k1 = s.fk + 1.0f;
    const float rk1 = rcp_fast(k1), rnk = rcp_fast(nk); // IEEE divisions cost ~10 instructions each here
    const float t1 = __fmul_rn(s.fm + 0.5f, log1pf(__fmul_rn(-d, rk1)));
    const float t2 = __fmul_rn(s.fn - s.fm + 0.5f, log1pf(__fmul_rn(d, rnk)));
    const float t3 = __fmul_rn(d, logf(__fmul_rn(r * nk, rk1)));
    // products and sums rounded one by one, in this order (no contraction: see the header note)
    float h = __fadd_rn(__fadd_rn(t1, t2), t3);
    h = __fadd_rn(h, stirling_tail(s.fm));
    h = __fadd_rn(h, stirling_tail(s.fn - s.fm));
    h = __fsub_rn(h, stirling_tail(s.fk));
    h = __fsub_rn(h, stirling_tail(s.fn - s.fk));
    return s.A <= h;
}

// Binomial(n, p), any p in [0, 1], lockstep form.  (bits_a, bits_b): random words for the first trial.
__device__ __forceinline__ int binomial(Rng &rng, int n, float p, uint32_t bits_a, uint32_t bits_b,
                                        const float *rcp_tab)
{
    if (n <= 0 || !(p > 0.0f))
        return 0;
    if (p >= 1.0f)
        return n;
    const bool flip = p > 0.5f;
    const float pp = flip ? 1.0f - p : p;
    DrawState s;
    int k;
    if ((float)n * pp < 10.0f) {
        inv_setup(s, n, pp);
        inv_start(s, bits_a);
        inv_run(s, rcp_tab);
        k = s.ix;
    } else {
        btrs_setup(s, n, pp);
        PairSource src(rng, bits_a, bits_b);
        while (true) {
            const int r = btrs_trial(s, src.a, src.b);
            if (r == 1 || (r == 2 && btrs_full(s)))
                break;
            src.advance();
        }
        k = (int)s.fk;
    }
    return flip ? n - k : k;
}

// ---- Poisson ---------------------------------------------------------------------------------
// lam is the reference's fp32 `hit` (sonics.pyx:419-420).  PTRS (Hoermann 1993) in the same
// split-fp32 style as BTRS: the proposal keeps the integer part of lam exactly, and the acceptance
// test uses  ln f(k) = k (log1p(y) - y) - ln(2 pi k)/2 - 1/(12k) + 1/(360 k^3),  y = (lam - k)/k,
// whose terms are O((k - lam)^2 / k) instead of O(lam ln lam).

// Poisson(lam) for 0 < lam < 10: sequential inversion, run to completion.
__device__ __forceinline__ int poisson_small(float lam, uint32_t bits)
{
    float px = __expf(-lam);
    float u = u01_fine(bits);
    int x = 0;
    while (u > px && x < 200) {
        ++x;
        u -= px;
        px *= lam * rcp_fast((float)x);
    }
    return x;
}

// PTRS state in DrawState: fn = lam, ci = floor(lam), cf = lam - ci + 0.43, a b vr, alpha = 1/alpha
__device__ __forceinline__ void pois_setup(DrawState &s, float lam) // lam >= 10
{
    s.fn = lam;
    const float slam = sqrt_fast(lam);
    s.b = fmaf(2.53f, slam, 0.931f);
    s.a = fmaf(0.02483f, s.b, -0.059f);
    s.alpha = fmaf(1.1328f, rcp_fast(s.b - 3.4f), 1.1239f); // 1 / alpha
    s.vr = fmaf(-3.6224f, rcp_fast(s.b - 2.0f), 0.9277f);
    s.ci = floorf(lam);
    s.cf = (lam - s.ci) + 0.43f;
    s.ix = 0;
}

// One PTRS trial: true = accepted (result s.ix)
__device__ __forceinline__ bool pois_trial(DrawState &s, uint32_t bits_a, uint32_t bits_b)
{
    const float u = u01_open(bits_a) - 0.5f;
    const float v = u01_open(bits_b);
    const float us = 0.5f - fabsf(u);
    const float rus = rcp_fast(us);
    const float fk = s.ci + floor_small(fmaf(fmaf(2.0f * s.a, rus, s.b), u, s.cf));
    s.ix = (int)fk;
    if (us >= 0.07f && v <= s.vr)
        return true;
    if (fk < 0.0f || (us < 0.013f && v > us))
        return false;
    const float lhs = log_fast(v * s.alpha * rcp_fast(fmaf(s.a * rus, rus, s.b)));
    float rhs;
    if (fk >= 10.0f) {
        const float rk = rcp_fast(fk);
        const float y = __fmul_rn(s.fn - fk, rk);
        const float t1 = __fmul_rn(fk, log1pf(y) - y);
        const float t3 = __fmul_rn(rk, fmaf(-0.0027777778f * rk, rk, 0.083333333f));
        rhs = __fsub_rn(fmaf(-0.5f, log_fast(6.2831853f * fk), t1), t3);
    } else {
        rhs = __fsub_rn(fmaf(fk, log_fast(s.fn), -s.fn), lgammaf(fk + 1.0f));
    }
    return lhs <= rhs;
}

// Poisson(lam), lockstep form.
__device__ __forceinline__ int poisson(Rng &rng, float lam)
{
    if (!(lam > 0.0f))
        return 0;
    const uint4 r0 = rng.block4();
    if (lam < 10.0f)
        return poisson_small(lam, r0.x);
    DrawState s;
    pois_setup(s, lam);
    PairSource src(rng, r0.x, r0.y);
    src.sa = r0.z;
    src.sb = r0.w;
    src.spare = 1;
    while (!pois_trial(s, src.a, src.b))
        src.advance();
    return s.ix;
}

} // namespace sonics
