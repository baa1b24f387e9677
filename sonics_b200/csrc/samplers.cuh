// Exact binomial / Poisson samplers for the PCR kernel.
//
// The reference calls numpy's legacy np.random.binomial (inversion / BTPE) at
// sonics.pyx:354,454,458,462,478 and np.random.poisson (PTRS) at :420.  Those
// are rejection / inversion samplers of the exact distributions; here the same
// distributions are sampled with algorithms that map better onto SIMT lanes:
//   binomial:  n*min(p,1-p) < 20  -> sequential inversion from 0 (BINV)
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
// Binomial(n, p): sequential inversion below this n min(p, q), BTRS from it on (BTRS needs >= 10; numpy switches at
// 30).  In lockstep a BTRS slot costs ~77 instructions of set-up + 2.6 rejection rounds of ~95, an inversion slot
// ~80 + 9 per step of its slowest lane: measured on B200 (profiles/r02_k1_ab.md) 10 -> 17 -> 20 -> 24 -> 30 -> 40 gives
// cfg2 1.089 -> 1.130 -> 1.133 -> 1.133 -> 1.131 -> 1.116e8 reactions/s (with the pinned histogram address).
#ifndef SONICS_INV_BELOW
#define SONICS_INV_BELOW 20.0f
#endif

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
__device__ __forceinline__ float log2_fast_raw(float x)
{
    float r;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
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
//   BTRS      fn p a b vr alpha ci cf fm, squeeze: sL sQ sC3 sC4 sR5 dmax | fk = candidate, A = ln of the scaled uniform
//   inversion a = (n+1) p/q, b = p/q         | fk = u (remaining uniform), A = px, ix = x, icap = cap
//   PTRS      fn = lam (the reference's fp32 `hit`), a b vr, alpha = 1/alpha | ix = candidate
struct DrawState {
    float fn, p, a, b, vr, alpha, ci, cf, fm, fk, A;
    float sL, sQ, sC3, sC4, sR5, dmax; // BTRS squeeze (btrs_setup)
    int ix, icap;
};

// ---- inversion: Binomial(n, p), p <= 0.5, n*p < SONICS_INV_BELOW ----
__device__ __forceinline__ void inv_setup(DrawState &v, int n, float p)
{
    const float q = 1.0f - p;
    v.b = p * rcp_fast(q);
    v.a = (float)(n + 1) * v.b;
    v.A = __expf((float)n * log1pf(-p)); // px = q^n
    v.icap = n < SONICS_RCP_TABLE - 1 ? n : SONICS_RCP_TABLE - 1; // P(X > 190 | np < 20) < 1e-100
    v.ix = 0;
    v.fk = 0.0f;
}
__device__ __forceinline__ void inv_start(DrawState &v, uint32_t bits) { v.fk = u01_fine(bits); }
// the sequential search, to its end (at most icap < SONICS_RCP_TABLE steps; n*p < 20: a handful typical):
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
// c0 = ln(p b0 / (q a0)) = log1p((frac - q) / (q a0)), frac = (n+1) p - m, as 2 atanh(w), w = eps / (2 + eps)
// (|eps| <= 1/(q a0) <= 0.18 for n p >= 10: five terms reach fp32 accuracy)
__device__ __forceinline__ float btrs_c0(float frac, float q, float a0)
{
    const float eps = __fmul_rn(frac - q, rcp_fast(__fmul_rn(q, a0)));
    const float w = __fmul_rn(eps, rcp_fast(2.0f + eps));
    const float w2 = __fmul_rn(w, w);
    float c0 = fmaf(w2, 0.11111111f, 0.14285714f);
    c0 = fmaf(c0, w2, 0.2f);
    c0 = fmaf(c0, w2, 0.33333333f);
    c0 = fmaf(c0, w2, 1.0f);
    return __fmul_rn(__fmul_rn(2.0f, w), c0);
}
__device__ __forceinline__ void btrs_setup(DrawState &s, int n, float p)
{
    s.fn = (float)n;
    s.p = p;
    const float q = 1.0f - p;
    // n*p exactly as hi + lo (n < 2^24, so fma recovers the rounding error of the product)
    const float hi = s.fn * p;
    const float lo = fmaf(s.fn, p, -hi);
    const float spq = sqrt_fast(hi * q);
    s.b = fmaf(2.53f, spq, 1.15f);
    const float rb = rcp_fast(s.b);
    s.a = fmaf(0.01f, p, fmaf(0.0248f, s.b, -0.0873f));
    s.vr = fmaf(-4.2f, rb, 0.92f);
    s.alpha = fmaf(5.1f, rb, 2.83f) * spq;
    s.ci = floorf(hi);                        // integer part of n*p (exact in fp32)
    s.cf = (hi - s.ci) + lo + 0.5f;           // c = n*p + 0.5 = ci + cf
    const float tf = s.cf - 0.5f + p;
    const float fl = floorf(tf);
    s.fm = s.ci + fl;                         // m = floor((n+1) p)
    // ---- squeeze: ln f(k)/f(m) as a quartic in d = k - m with a rigorous remainder ----
    // With a0 = m+1, b0 = n-m+1, ra = 1/a0, rb = 1/b0 the deviance form of btrs_full() expands to
    //   ln f(k)/f(m) = L d - Q d^2 + C3 d^3 + C4 d^4 + R,   |R| <= 0.11 |d|^5 (ra^4 + rb^4)  for |d| <= 0.9 min(a0, b0)
    //   L  = (ra - rb)/2 + c0 + (ra^2 - rb^2)/12            c0 = ln(p b0 / (q a0)), see btrs_full()
    //   Q  = (ra + rb)/2 + (ra^2 + rb^2)/4 + (ra^3 + rb^3)/12
    //   C3 = (ra^2 - rb^2)/6 + (ra^3 - rb^3)/6 + (ra^4 - rb^4)/12
    //   C4 = -(ra^3 + rb^3)/12 - (ra^4 + rb^4)/8
    // (-g(a;a0) - g(b;b0), ln(ab/(a0 b0))/2 and the Stirling tails, each expanded in d/a0 and d/b0; the remainder
    // constant is checked against the exact log-pmf ratio for 10 <= np, n <= 5e6 in tests/test_sampler_math.py,
    // worst observed 0.11 -> 0.25 used).  BTPE's squeeze (K&S step 5.2) is second order: its band ~|d|/(2npq) left
    // 5 % of the trials undecided at npq = 1000 and 37 % at npq = 10; this one 0.2 % and 7 %, so the ~140-instruction
    // exact test, which a warp executes for its 1-3 undecided lanes, runs in few of the rounds instead of most.
    const float a0 = s.fm + 1.0f, b0 = s.fn - s.fm + 1.0f;
    const float ra = rcp_fast(a0), rbb = rcp_fast(b0);
    const float ra2 = __fmul_rn(ra, ra), rb2 = __fmul_rn(rbb, rbb);
    const float ra3 = __fmul_rn(ra2, ra), rb3 = __fmul_rn(rb2, rbb);
    const float ra4 = __fmul_rn(ra2, ra2), rb4 = __fmul_rn(rb2, rb2);
    const float c0 = btrs_c0(tf - fl, q, a0);
    s.sL = fmaf(0.083333333f, ra2 - rb2, fmaf(0.5f, ra - rbb, c0));
    s.sQ = fmaf(0.083333333f, ra3 + rb3, fmaf(0.25f, ra2 + rb2, __fmul_rn(0.5f, ra + rbb)));
    s.sC3 = fmaf(0.083333333f, ra4 - rb4, __fmul_rn(0.16666667f, (ra2 - rb2) + (ra3 - rb3)));
    s.sC4 = fmaf(-0.125f, ra4 + rb4, __fmul_rn(-0.083333333f, ra3 + rb3));
    s.sR5 = __fmul_rn(0.25f, ra4 + rb4);
    s.dmax = floorf(__fmul_rn(0.9f, fminf(a0, b0)));
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
    // quartic squeeze (btrs_setup).  Beyond dmax the pmf is below its value at m +- dmax (unimodal), so the upper
    // bound at the clamped distance still rejects; acceptance needs the candidate in range.
    const float d = s.fk - s.fm;
    const float dc = fminf(fmaxf(d, -s.dmax), s.dmax);
    const float t = __fmul_rn(dc, fmaf(dc, fmaf(dc, fmaf(dc, s.sC4, s.sC3), -s.sQ), s.sL));
    const float d2 = __fmul_rn(dc, dc);
    const float rho = __fmul_rn(__fmul_rn(__fmul_rn(d2, d2), fabsf(dc)), s.sR5);
    const float eps = fmaf(1e-5f, fabsf(t), 1e-5f); // margin for the fp32 evaluation of A, t and rho
    if (s.A > __fadd_rn(__fadd_rn(t, rho), eps))
        return 0;
    if (d == dc && s.A < __fsub_rn(__fsub_rn(t, rho), eps))
        return 1;
    return 2;
}

// The exact test for an undecided candidate: A <= ln f(k)/f(m).
//
// With a = k+1, b = n-k+1, a0 = m+1, b0 = n-m+1, d = k-m and Stirling's series in Hoermann's (k+1) form,
//   ln f(k)/f(m) = -g(a; a0) - g(b; b0) + 1/2 ln(a b / (a0 b0)) + d c0 + fc(m) + fc(n-m) - fc(k) - fc(n-k)
//   g(x; x0) = x ln(x/x0) - (x - x0)               the deviance term, O(d^2 / x0) instead of O(d)
//   c0       = ln(p b0 / (q a0)) = log1p((frac - q) / (q a0)),  frac = (n+1) p - m      a per-draw constant
// (the saddle-point arrangement of Loader, "Fast and accurate computation of binomial probabilities", 2000).  The
// textbook form sums three O(d) logarithm terms that cancel to O(d^2/npq): in fp32 it is off by up to 3e-3 at
// n = 5e6 and needs two log1pf and one logf (~350 instructions executed by the 2-3 lanes of a warp that are
// undecided).  Here every term is small by itself: g through x ln(x/x0) = 2 x atanh(v), v = d / (x + x0), as the
// series d v + 2 x (v^3/3 + v^5/5 + ...), c0 the same way; |error| <= ~1e-5 for n up to 5e6 (CPU test in
// tests/test_sampler_math.py), ~90 instructions without a slow-path call.  Candidates with |v| > 1/2 (k + 1 beyond
// 3 (m + 1) or below (m + 1) / 3: far tails of the smallest means) take the logarithm directly.
__device__ __forceinline__ float atanh_tail(float v) // v^3/3 + v^5/5 + ... + v^19/19
{
    const float v2 = __fmul_rn(v, v);
    float s = 0.052631579f;
    s = fmaf(s, v2, 0.058823529f);
    s = fmaf(s, v2, 0.066666667f);
    s = fmaf(s, v2, 0.076923077f);
    s = fmaf(s, v2, 0.090909091f);
    s = fmaf(s, v2, 0.11111111f);
    s = fmaf(s, v2, 0.14285714f);
    s = fmaf(s, v2, 0.2f);
    s = fmaf(s, v2, 0.33333333f);
    return __fmul_rn(__fmul_rn(s, v2), v);
}
// g(x; x0) = x ln(x / x0) - dx,  dx = x - x0
__device__ __forceinline__ float deviance_term(float x, float x0, float dx)
{
    const float v = __fmul_rn(dx, rcp_fast(x + x0));
    if (fabsf(v) > 0.5f)
        return fmaf(x, logf(__fdiv_rn(x, x0)), -dx);
    return fmaf(__fmul_rn(2.0f, x), atanh_tail(v), __fmul_rn(dx, v));
}
__device__ __forceinline__ bool btrs_full(const DrawState &s)
{
    const float q = 1.0f - s.p;
    const float d = s.fk - s.fm;
    const float a0 = s.fm + 1.0f, b0 = s.fn - s.fm + 1.0f;
    const float a = s.fk + 1.0f, b = s.fn - s.fk + 1.0f;
    const float ga = deviance_term(a, a0, d);
    const float gb = deviance_term(b, b0, -d);
    const float lg = __fmul_rn(0.5f, log_fast(__fmul_rn(__fmul_rn(a, rcp_fast(a0)), __fmul_rn(b, rcp_fast(b0)))));
    // c0: frac = (n+1) p - m from the split product of btrs_setup (cf = frac(n p) + 1/2, fm - ci = floor(...))
    const float c0 = btrs_c0((s.cf - 0.5f + s.p) - (s.fm - s.ci), q, a0);
    // products and sums rounded one by one, in this order (no contraction: see the header note)
    float h = __fsub_rn(lg, __fadd_rn(ga, gb));
    h = fmaf(d, c0, h);
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
    if ((float)n * pp < SONICS_INV_BELOW) {
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

// ---- Binomial(n, p), p <= 0.5, n p q large: the opt-in approximate sampler (sonics_set_samplers) ------------
// One normal deviate (Box-Muller on one 64-bit block) pushed through the first Cornish-Fisher term and rounded:
//     x = n p + sqrt(n p q - 1/12) z + (q - p) (z^2 - 1) / 6,      k = floor(x + 1/2) clipped to [0, n]
// (the skewness term; -1/12 is Sheppard's correction for the rounding).  Not exact: against the binomial cdf the
// error is 1.2e-2 / (n p q) at most (6e-4 at n p q = 20, 1.2e-4 at 100, 1.2e-5 at 1000), the pmf is within
// 0.8 % / 0.07 % inside three sigma at n p q = 100 / 1000 -- computed exactly in tests/test_sampler_math.py, sampled on
// the device in tests/test_gpu_fast_samplers.py.  No rejection: every lane of a warp finishes in the same ~25
// instructions after its Philox block, where BTRS in lockstep needs 2.6 rounds.
__device__ __forceinline__ int binomial_normal(int n, float p, uint32_t bits_a, uint32_t bits_b)
{
    const float fn = (float)n, q = 1.0f - p;
    const float mu = fn * p;
    const float s = sqrt_fast(fmaf(mu, q, -0.083333333f));
    const float u1 = fmaf(__uint2float_rz(bits_a), 2.3283064365386963e-10f, 1.1641532182693481e-10f); // (0, 1)
    const float r = sqrt_fast(-1.3862943611f * log2_fast_raw(u1)); // sqrt(-2 ln u1)
    const float z = r * __cosf(6.2831853072f * u01_open(bits_b));
    const float x = fmaf(s, z, mu) + (q - p) * 0.16666667f * fmaf(z, z, -1.0f);
    const float k = floorf(x + 0.5f);
    return (int)fminf(fmaxf(k, 0.0f), fn);
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
