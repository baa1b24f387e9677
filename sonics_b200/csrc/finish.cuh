// End of a simulation: lnL of the final pool and, on request, the sequencing readout with its descriptors.
// Shared by K1b (k_lnl, one lane per pool column of the K1a scratch) and by the fused lockstep kernel
// (k_pcr_ls, the column is the lane's histogram in shared memory): both layouts keep bin c of a simulation at
// col[c * 32].
//
//   lnL       sonics.pyx:360-363 -> pymc_extracted.f:63-109 (lnl.cuh)
//   readout   sonics.pyx:341-357 (n_i ~ Binomial(D, pool_i / T) per bin), :365-372 (choice(mid, D) + bincount)
//   identity  sonics.pyx:373-383
//   r_squared rsq(), sonics.pyx:273-291, stored through a C float (:300)
#pragma once
#include <stdint.h>

#include "lnl.cuh"
#include "rng.cuh"
#include "samplers.cuh"

namespace sonics {

struct GtView {
    int W, lo, A, D;
    const int32_t *al;  // non-zero input alleles, ascending
    const int32_t *rd;  // their reads
    const double *fl;   // factln(reads)
    double fl_D;        // factln(D)
};

// lnL of the pool in `col` against the genotype's reads; T = sum(pool) is returned for the readout.
__device__ __forceinline__ double pool_lnl(const int *col, const GtView &g, int &T_out)
{
    // sentinel: any reads[i] > pool[i] (:360-361); T = sum(pool)
    bool sentinel = false;
    for (int i = 0; i < g.A; ++i)
        sentinel |= g.rd[i] > col[(g.al[i] - g.lo) * 32];
    int T = 0;
    for (int c = 0; c < g.W; ++c)
        T += col[c * 32];
    T_out = T;
    if (sentinel)
        return SONICS_SENTINEL;
    double like = 0.0;
    int ia = 0;
    for (int c = 0; c < g.W; ++c) {
        const int cnt = col[c * 32];
        int xr = 0;
        double flx = 0.0;
        if (ia < g.A && g.al[ia] - g.lo == c) {
            xr = g.rd[ia];
            flx = g.fl[ia];
            ++ia;
        }
        if (cnt == 0)
            continue; // x == 0 too: the reference adds exact zeros
        const double fc = factln_dev(cnt);
        like = __dadd_rn(like, fc);
        like = __dsub_rn(like, flx); // factln(0) = gammln(1) = 0.0 exactly in the reference
        like = __dsub_rn(like, xr ? factln_dev(cnt - xr) : fc);
    }
    return __dsub_rn(like, __dsub_rn(__dsub_rn(factln_dev(T), g.fl_D), factln_dev(T - g.D)));
}

// Sequencing readout of the pool in `col` (overwritten: first with the stage-1 counts n_i, then with the readout).
// rr: the simulation's STREAM_READOUT generator.  dump (nullable): the readout per window bin, dump[c], c < W.
__device__ __forceinline__ void pool_readout(int *col, const GtView &g, int T, Rng &rr, const float *rcp_tab,
                                             double &ident_out, float &r2_out, int32_t *dump)
{
    const int W = g.W, A = g.A, D = g.D, lo = g.lo;
    // stage 1: n_i ~ Binomial(D, pool_i / T) for every bin (:354); the column is reused as storage
    int N = 0;
    const double invT = T > 0 ? 1.0 / (double)T : 0.0;
    for (int c = 0; c < W; ++c) {
        const int cnt = col[c * 32];
        int ni = 0;
        if (cnt != 0) {
            const uint4 w = rr.block4();
            ni = binomial(rr, D, (float)((double)cnt * invT), w.x, w.y, rcp_tab);
        }
        col[c * 32] = ni;
        N += ni;
    }
    double ident = 0.0;
    float r2 = -999999.0f;
    if (N > 0) {
        // stage 2: readout = bincount(choice(mid, D)) ~ Multinomial(D; n_i / N) (:368),
        // drawn as a chain of conditional binomials
        int R = D, M = N, fr = -1, to = -1, acc = 0, ia = 0;
        for (int c = 0; c < W; ++c) {
            const int ni = col[c * 32];
            int xi = 0;
            if (ni > 0 && R > 0) {
                const uint4 w = rr.block4();
                xi = (ni == M) ? R : binomial(rr, R, __fdividef((float)ni, (float)M), w.x, w.y, rcp_tab);
            }
            R -= xi;
            M -= ni;
            col[c * 32] = xi;
            int xr = 0;
            if (ia < A && g.al[ia] - lo == c) {
                xr = g.rd[ia];
                ++ia;
            }
            acc += min(xr, xi);
            if (xr != 0 || xi != 0) {
                if (fr < 0)
                    fr = c;
                to = c;
            }
        }
        ident = (double)acc / (double)D;
        const double mean = (double)D / (double)(to - fr + 1);
        double ss_tot = 0.0;
        long long ss_res = 0;
        ia = 0;
        while (ia < A && g.al[ia] - lo < fr)
            ++ia;
        for (int c = fr; c <= to; ++c) {
            int xr = 0;
            if (ia < A && g.al[ia] - lo == c) {
                xr = g.rd[ia];
                ++ia;
            }
            const int xi = col[c * 32];
            const double dl = (double)xr - mean;
            ss_tot = __dadd_rn(ss_tot, __dmul_rn(dl, dl));
            const long long e = (long long)xr - xi;
            ss_res += e * e;
        }
        if (ss_tot == 0.0)
            ss_tot = 1e-16;
        r2 = (float)(1.0 - (double)ss_res / ss_tot);
    } // else nothing sequenced (:369-372): the column already holds the all-zero readout
    if (dump)
        for (int c = 0; c < W; ++c)
            dump[c] = col[c * 32];
    ident_out = ident;
    r2_out = r2;
}

} // namespace sonics
