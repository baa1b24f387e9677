// K1 (lockstep form): 32 SIMILAR simulations per warp, every loop bound warp-uniform, lnL fused.
//
// Same simulation as k_pcr (sim_kernel.cuh) -- same draws in the same order from the same Philox streams, hence
// bit-identical lnL / group / identity / r^2 (tests/test_gpu_lockstep.py) -- scheduled differently:
//
//   k_sort_keys  one key per work unit from the simulation's priors (which depend on (seed, uid, j) only):
//                noise flag | start pair (absolute repeat counts) | reach of the noise | Morton(eff, down, up)
//                -- see SortCfg below
//   CUB radix sort of (key, unit)
//   k_pcr_ls     warp w simulates 32 consecutive units of the sorted list in lockstep over
//                (cycle, bin, draw stage).
//
// Why: in k_pcr lanes drift apart (different start alleles, efficiencies, rejection counts), half of the executed
// instructions are the per-lane state machine, and the samplers run at 3-16 of 32 lanes (profiles/r01_k1a).  Lanes
// that hold the same start alleles (of any genotype), the same noise reach and neighbouring prior parameters walk the same bins with
// counts that differ by O(sqrt(n)), so in lockstep they take the same sampler (inversion / BTRS) at the same
// (cycle, bin, stage) slot, the bin walk and the cycle boundary cost a few warp-uniform instructions per visit, and
// all 32 lanes finish together, so the lnL (and the readout) runs fully converged inside the kernel: no pool
// hand-off through HBM.  What remains divergent is what is random per draw: rejected trials and the Stirling test.
//
// A lane only touches bins [sb, bend] of its own sweep (exactly the bins k_pcr visits) and the warp walks absolute
// repeat counts, so warps that mix genotypes (different windows, reads, allele lists) are the normal case.
#pragma once
#include <stdint.h>

#include "sim_kernel.cuh"

namespace sonics {

#ifndef SONICS_LS_MIN_BLOCKS
#define SONICS_LS_MIN_BLOCKS 6
#endif

// A rejection sampler accepts with probability > 0.6 per trial; this bound only keeps a corrupted operand (NaN)
// from hanging the warp.
#define SONICS_LS_MAX_TRIALS 4096

// Diagnostics build only (-DSONICS_LS_STATS, tools/ls_stats.py): warp-level counters of the lockstep schedule.
#ifdef SONICS_LS_STATS
__device__ unsigned long long g_ls_stats[128];
__shared__ unsigned int s_ls_stats[4][128]; // per warp, written by lane 0, flushed when the warp leaves the kernel
#define LS_STAT(i, v)                                                                                                  \
    do {                                                                                                               \
        const unsigned int ls_v_ = (unsigned int)(v); /* warp-wide votes in v run with all lanes */                    \
        if ((threadIdx.x & 31) == 0)                                                                                   \
            s_ls_stats[threadIdx.x >> 5][i] += ls_v_;                                                                  \
    } while (0)
#else
#define LS_STAT(i, v)                                                                                                  \
    do {                                                                                                               \
    } while (0)
#endif

// Sort key of a work unit.  What a reaction does depends on its start alleles, the noise molecules and its priors --
// the genotype's reads only enter the lnL -- so like simulations are pooled across ALL genotypes of a launch (a
// genotyping round runs 100-400 simulations per genotype, 12-50 per start-allele group: less than a warp):
//   pair_major = 1 (default)
//     noise flag | lower, higher start allele (absolute repeat counts) | reach of the noise below / above them | priors
//     Noise molecules sit on every input allele (sonics.pyx:334-337): with them the live bins of a reaction are the
//     genotype's whole input range; 85 % of the genotypes of a lobSTR-style VCF get them from the automatic rule
//     (sonics:407-416).  The three priors are interleaved bit by bit (Morton order), so any run of 32 neighbours of
//     the sorted list is as compact in (eff, down, up) as the population of its cell allows, whatever that is.
//     Measured (B200, profiles/r02_k1_ab.md): cfg4 genotyping 114 k -> 152 k genotypes/s against the genotype-major
//     key, single-locus launches +4.5 % against eff | down | up bit fields.
//   pair_major = 0 (sonics_set_sort_mode(SONICS_SORT_SLOT_MAJOR))
//     genotype slot | start-allele group | eff | down | up as bit fields (sonics_set_sort_bits).
// The fields are as wide as the table needs (lowest input allele and widest input range of its genotypes, noise
// present or not), so a single locus sorts 32-bit keys in three radix passes.
struct SortCfg {
    int grp_bits, eff_bits, dn_bits, up_bits; // slot-major key
    int pair_major;
    int al_base, al_bits, span_bits, noise_bits, morton_bits; // pair-major key
};
#define SONICS_LS_MORTON_BITS 5 // per prior

template <typename KeyT>
__global__ void k_sort_keys(const SimArgs a, const SortCfg c, KeyT *keys, uint32_t *vals)
{
    const int u = blockIdx.x * blockDim.x + threadIdx.x;
    if (u >= a.n_units)
        return;
    int slot, jrel;
    unit_decode(a, u, slot, jrel);
    const int g = a.active ? a.active[slot] : slot;
    const GtHeader h = a.t.hdr[g];
    int64_t j = a.sim_begin + jrel;
    if (a.replay) {
        j = a.replay[g].best_sim;
        if (j < 0)
            j = 0;
    }
    const Priors pr = draw_priors(a.p, a.seed, h, j);
    // start-allele group: the second allele's index (the first is fixed, :323), both with --random (:321)
    const uint32_t gid = a.p.random_start ? (uint32_t)(min(pr.i1, pr.i2) * h.n_alleles + max(pr.i1, pr.i2))
                                          : (uint32_t)pr.i2;
    auto quant = [](double v, double lo, double rng, int bits) -> uint32_t {
        if (bits <= 0 || !(rng > 0.0))
            return 0u;
        const double t = (v - lo) / rng;
        const int q = (int)(t * (double)(1 << bits));
        return (uint32_t)max(0, min(q, (1 << bits) - 1));
    };
    const double up_rng = a.p.up_preference ? a.p.up_hi - a.p.up_lo : a.p.down_lo + a.p.down_rng - a.p.up_lo;
    KeyT k;
    if (c.pair_major) {
        const int v1 = a.t.allele[h.a_off + pr.i1], v2 = a.t.allele[h.a_off + pr.i2];
        const int vlo = min(v1, v2), vhi = max(v1, v2);
        const bool noisy = h.noise_add > 0;
        // span_bits hold 0 .. the widest input range of the table: start pair distance and both reaches fit
        k = (KeyT)(noisy && c.noise_bits ? 1 : 0);
        k = (k << c.al_bits) | (KeyT)(vlo - c.al_base);
        k = (k << c.span_bits) | (KeyT)(vhi - vlo);
        if (c.noise_bits) {
            const int below = noisy ? vlo - (h.lo + h.in_lo_bin) : 0;
            const int above = noisy ? a.t.allele[h.a_off + h.n_alleles - 1] - vhi : 0;
            k = (k << c.span_bits) | (KeyT)below;
            k = (k << c.span_bits) | (KeyT)above;
        }
        const uint32_t qe = quant(pr.eff, a.p.eff_lo, a.p.eff_rng, c.morton_bits);
        const uint32_t qd = quant(pr.down, a.p.down_lo, a.p.down_rng, c.morton_bits);
        const uint32_t qu = quant(pr.up, a.p.up_lo, up_rng, c.morton_bits);
        for (int i = c.morton_bits - 1; i >= 0; --i)
            k = (k << 3) | (KeyT)((((qe >> i) & 1u) << 2) | (((qd >> i) & 1u) << 1) | ((qu >> i) & 1u));
    } else {
        const uint32_t qe = quant(pr.eff, a.p.eff_lo, a.p.eff_rng, c.eff_bits);
        const uint32_t qd = quant(pr.down, a.p.down_lo, a.p.down_rng, c.dn_bits);
        const uint32_t qu = quant(pr.up, a.p.up_lo, up_rng, c.up_bits);
        k = (KeyT)slot;
        k = (k << c.grp_bits) | (KeyT)(gid & ((1u << c.grp_bits) - 1u));
        k = (k << c.eff_bits) | (KeyT)qe;
        k = (k << c.dn_bits) | (KeyT)qd;
        k = (k << c.up_bits) | (KeyT)qu;
    }
    keys[u] = k;
    vals[u] = (uint32_t)u;
}

// ---- lockstep samplers: every lane of the warp calls these together ---------------------------------------
// Same pieces, same random words per draw as the state machine of k_pcr: an inversion takes one pair of its
// simulation's stream, a rejection sampler one pair per trial, the Stirling test none.
// FAST (sonics_set_samplers): draws with n p q >= fast_npq take binomial_normal() instead of BTRS -- one block, no
// rejection.  The exact instantiation (FAST = false) contains none of it.
template <bool FAST>
__device__ __forceinline__ int binomial_ls(PairRng &rng, int n, float p, const float *rcp_tab, float fast_npq)
{
    int x = 0;
    bool need = n > 0 && p > 0.0f; // !(p > 0): degenerate priors only
    if (need && p >= 1.0f) {
        x = n;
        need = false;
    }
    const bool flip = p > 0.5f;
    const float pp = flip ? 1.0f - p : p;
    const bool inv = need && (float)n * pp < SONICS_INV_BELOW;
    const bool fast = FAST && need && !inv && (float)n * pp * (1.0f - pp) >= fast_npq;
    bool pend = need && !inv && !fast;
    if (FAST && __any_sync(SONICS_FULL, fast)) {
        if (fast) {
            const uint2 w = rng.pair();
            const int k = binomial_normal(n, pp, w.x, w.y);
            x = flip ? n - k : k;
        }
    }
    int guard = 0;
    DrawState ds;
#ifdef SONICS_LS_STATS
    {
        const unsigned mp = __ballot_sync(SONICS_FULL, pend), mi = __ballot_sync(SONICS_FULL, inv);
        LS_STAT(1, 1);
        LS_STAT(2, mp != 0);
        LS_STAT(3, mi != 0);
        LS_STAT(4, mp != 0 && mi != 0);
        LS_STAT(5, __popc(mp));
        LS_STAT(6, __popc(mi));
        // lanes by log2(n p q) of their BTRS draw
        const float npq = (float)n * pp * (1.0f - pp);
        const int bk = pend ? min(15, max(0, (int)floorf(log2f(npq)))) : -1;
        for (int k = 3; k < 16; ++k)
            LS_STAT(32 + k, __popc(__ballot_sync(SONICS_FULL, bk == k)));
    }
    int rounds = 0;
#endif
    if (__any_sync(SONICS_FULL, pend)) {
        if (pend)
            btrs_setup(ds, n, pp);
        // The exact test of an undecided candidate runs inside the round (deferring it to a common point was
        // measured: fewer executions, but the lanes it sends back retry alone -- 7.85e7 vs 8.46e7 reactions/s).
        do {
#ifdef SONICS_LS_STATS
            ++rounds;
            LS_STAT(7, 1);
            LS_STAT(8, __popc(__ballot_sync(SONICS_FULL, pend)));
            bool und = false;
#endif
            if (pend) {
                const uint2 w = rng.pair();
                int r = btrs_trial(ds, w.x, w.y);
#ifdef SONICS_LS_STATS
                und = r == 2;
#endif
                if (r == 2)
                    r = btrs_full(ds) ? 1 : 0;
                if (r == 1) {
                    const int k = (int)ds.fk;
                    x = flip ? n - k : k;
                    pend = false;
                }
            }
#ifdef SONICS_LS_STATS
            {
                const unsigned mu = __ballot_sync(SONICS_FULL, und);
                LS_STAT(9, mu != 0);
                LS_STAT(10, __popc(mu));
            }
#endif
        } while (__any_sync(SONICS_FULL, pend) && ++guard < SONICS_LS_MAX_TRIALS);
#ifdef SONICS_LS_STATS
        LS_STAT(16 + min(rounds, 15), 1);
#endif
    }
    if (__any_sync(SONICS_FULL, inv)) {
        if (inv) {
            inv_setup(ds, n, pp);
            const uint2 w = rng.pair();
            inv_start(ds, w.x);
            inv_run(ds, rcp_tab);
            x = flip ? n - ds.ix : ds.ix;
        }
#ifdef SONICS_LS_STATS
        {
            const int steps = inv ? ds.ix : 0;
            LS_STAT(11, __reduce_max_sync(SONICS_FULL, steps));
            LS_STAT(12, __reduce_add_sync(SONICS_FULL, steps));
        }
#endif
    }
    return x;
}

__device__ __forceinline__ int poisson_ls(PairRng &rng, bool need, float lam)
{
    int x = 0;
    const bool small = need && lam < 10.0f;
    bool pend = need && !small;
    int guard = 0;
    DrawState ds;
    if (__any_sync(SONICS_FULL, pend)) {
        if (pend)
            pois_setup(ds, lam);
        do {
            if (pend) {
                const uint2 w = rng.pair();
                if (pois_trial(ds, w.x, w.y)) {
                    x = ds.ix;
                    pend = false;
                }
            }
        } while (__any_sync(SONICS_FULL, pend) && ++guard < SONICS_LS_MAX_TRIALS);
    }
    if (__any_sync(SONICS_FULL, small)) {
        if (small) {
            const uint2 w = rng.pair();
            x = poisson_small(lam, w.x);
        }
    }
    return x;
}

// Shared memory: [rcp table][warp 0 hist: Wcap x 32 int32][warp 1 hist]...   (same as k_pcr)
__device__ __forceinline__ int lds_u32(uint32_t addr)
{
    int v;
    asm volatile("ld.shared.s32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts_u32(uint32_t addr, int v)
{
    asm volatile("st.shared.s32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}

template <bool READOUT, bool FAST>
__global__ void __launch_bounds__(128, SONICS_LS_MIN_BLOCKS) k_pcr_ls(const SimArgs a)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int Wcap = a.Wcap;
    float *rcp_tab = reinterpret_cast<float *>(smem_raw);
    for (int i = threadIdx.x; i < SONICS_RCP_TABLE; i += blockDim.x)
        rcp_tab[i] = i ? 1.0f / (float)i : 0.0f;
#ifdef SONICS_LS_STATS
    for (int i = lane; i < 128; i += 32)
        s_ls_stats[warp][i] = 0;
#endif
    __syncthreads();
    int *hist = reinterpret_cast<int *>(smem_raw + SONICS_RCP_TABLE * 4 + (size_t)warp * warp_smem_bytes(Wcap)) + lane;
    // the sweep addresses its column through one 32-bit shared-memory address that the compiler cannot rebuild from
    // threadIdx at every bin (it did: 11 instructions per sweep step; +1.2 % on cfg2)
    uint32_t hs;
    asm volatile("mov.u32 %0, %1;" : "=r"(hs) : "r"((uint32_t)__cvta_generic_to_shared(hist)));
#define LS_LD(bin) lds_u32(hs + 128u * (uint32_t)(bin))
#define LS_ST(bin, val) sts_u32(hs + 128u * (uint32_t)(bin), (val))
    const int n_cycles = a.p.n_cycles, cap_cycle = a.p.capture_cycle;
    const int BIG = 1 << 29;

    while (true) {
        // next batch of 32 units of the sorted list
        unsigned int batch = 0;
        if (lane == 0)
            batch = atomicAdd(a.counter, 1u);
        batch = __shfl_sync(SONICS_FULL, batch, 0);
        if ((int64_t)batch * 32 >= (int64_t)a.n_units)
            break;
        const int idx = (int)batch * 32 + lane;
        bool valid = idx < a.n_units;
        LS_STAT(0, 1);

        // ---- per-lane set-up: priors, start alleles, start pool (sonics.pyx:302-337) ----
        int u = 0, g = 0, jrel = 0, W = 0, lo = 0, floor_bin = 0;
        int64_t j = 0;
        float eff = 0.f, capture = 0.f, l_up = 0.f, l_dn = 0.f, e1_up = 0.f, e1_dn = 0.f;
        int live_lo = BIG, live_hi = -1;
        int b1 = 0, b2 = 0, noise_add = 0;
        PairRng rng;
        rng.init(a.seed, 0, 0);
        if (valid) {
            u = a.order ? (int)a.order[idx] : idx;
            int slot;
            unit_decode(a, u, slot, jrel);
            g = a.active ? a.active[slot] : slot;
            j = a.sim_begin + jrel;
            if (a.replay) {
                j = a.replay[g].best_sim;
                valid = j >= 0; // no_success: nothing to replay
            }
        }
        if (valid) {
            const GtHeader h = a.t.hdr[g];
            W = h.W;
            lo = h.lo;
            floor_bin = h.floor_al - lo;
            noise_add = h.noise_add;
            const int A = h.n_alleles;
            const Priors pr = draw_priors(a.p, a.seed, h, j);
            rng.init(a.seed, (uint64_t)j, h.uid); // the draws of the PCR cycles
            eff = (float)pr.eff; // cdef float efficiency, capture (:401)
            capture = (float)pr.cap;
            l_up = log1pf(-(float)pr.up);   // ln(1 - up)
            l_dn = log1pf(-(float)pr.down); // ln(1 - down)
            e1_up = expm1f(l_up);           // per-bin step of expm1(al * L)
            e1_dn = expm1f(l_dn);
            b1 = a.t.allele[h.a_off + pr.i1] - lo;
            b2 = a.t.allele[h.a_off + pr.i2] - lo;
            live_lo = min(b1, b2);
            live_hi = max(b1, b2);
            if (!a.replay) {
                const int64_t o = (int64_t)g * a.out_stride + a.out_offset + jrel;
                a.group[o] = (uint16_t)(min(pr.i1, pr.i2) * A + max(pr.i1, pr.i2));
                if (a.simpar) {
                    a.simpar[o * 4 + 0] = pr.down;
                    a.simpar[o * 4 + 1] = pr.up;
                    a.simpar[o * 4 + 2] = pr.cap;
                    a.simpar[o * 4 + 3] = pr.eff;
                }
            }
        }
        const int Wmax = __reduce_max_sync(SONICS_FULL, W);
        for (int c = 0; c < Wmax; ++c)
            hist[c * 32] = 0;
        if (valid) {
            hist[b1 * 32] += a.p.half_copies;
            hist[b2 * 32] += a.p.half_copies;
            if (noise_add > 0) { // :334-337
                const GtHeader h = a.t.hdr[g];
                for (int i = 0; i < h.n_alleles; ++i)
                    hist[(a.t.allele[h.a_off + i] - lo) * 32] += noise_add;
                live_lo = h.in_lo_bin;
                live_hi = a.t.allele[h.a_off + h.n_alleles - 1] - lo;
            }
        }

        // ---- simulate (:397-480) ----
        int ct = 0; // function-scope `ct` of the reference; goes stale below the floor (:478)
        for (int cycle = 1; cycle <= n_cycles; ++cycle) {
            if (cycle == cap_cycle) {
                // capture step (:412-423): max count, first non-zero allele, non-zero count; then one Poisson
                // draw per non-zero bin
                // the warp walks ABSOLUTE repeat counts (lanes of different genotypes have different window origins)
                const int wl = __reduce_min_sync(SONICS_FULL, lo + live_lo);
                const int wh = __reduce_max_sync(SONICS_FULL, lo + live_hi);
                int maxp = 0, nnz = 0, minidx = -1;
                for (int cw = wl; cw <= wh; ++cw) {
                    const int c = cw - lo;
                    if (c >= live_lo && c <= live_hi) {
                        const int v = hist[c * 32];
                        maxp = max(maxp, v);
                        if (v != 0) {
                            ++nnz;
                            if (minidx < 0)
                                minidx = lo + c;
                        }
                    }
                }
                float cap_set = 0.0f, floor_cap = 0.0f;
                if (nnz > 0) {
                    cap_set = (float)((double)capture / (double)(maxp - minidx));
                    floor_cap = (float)(1.0 - ((double)__fmul_rn(cap_set, (float)nnz)) / 2.0);
                }
                int cap_n = 1;
                for (int cw = wl; cw <= wh; ++cw) {
                    const int c = cw - lo;
                    const int v = (nnz > 0 && c >= live_lo && c <= live_hi) ? hist[c * 32] : 0;
                    float lam = 0.0f;
                    if (v != 0) {
                        ct = v;
                        lam = __fmul_rn(__fadd_rn(floor_cap, __fmul_rn(cap_set, (float)cap_n)), (float)v);
                    }
                    // numpy raises for lam < 0; unreachable with sane ranges: treated as 0
                    const int x = poisson_ls(rng, v != 0 && lam > 0.0f, lam);
                    if (v != 0) {
                        hist[c * 32] = x < v ? x : v; // ct_up = min(poisson(hit), ct) (:421)
                        ++cap_n;
                    }
                }
            }
            // ---- amplification sweep of this cycle (:424-479) ----
            const int sb = live_lo;
            const int bend = valid ? min(live_hi + 1, W - 1) : -1;
            const int wlo = __reduce_min_sync(SONICS_FULL, lo + sb);
            const int wend = __reduce_max_sync(SONICS_FULL, lo + bend);
            int carry = 0;
            int prev = (valid && sb > 0) ? LS_LD(sb - 1) : 0; // finished value of bin b-1
            float E_up = 0.0f, E_dn = 0.0f;
            if (valid) {
                const float al = (float)(lo + sb);
                E_up = expm1f(__fmul_rn(al, l_up));
                E_dn = expm1f(__fmul_rn(al, l_dn));
            }
            for (int bw = wlo; bw <= wend; ++bw) {
                const int b = bw - lo;
                const bool in = b >= sb && b <= bend;
                LS_STAT(13, 1);
                LS_STAT(14, __popc(__ballot_sync(SONICS_FULL, in)));
                int cur = 0, n0 = 0;
                bool stutter = false;
                float p_slip = 0.0f, p_upn = 0.0f;
                if (in) {
                    const int raw = LS_LD(b);
                    cur = raw + carry;
                    carry = 0;
                    if (raw != 0) { // empty at cycle start: not amplified this cycle (:424)
                        stutter = b >= floor_bin;
                        if (stutter) {
                            ct = cur;
                            const float prob_up = -E_up, prob_down = -E_dn;
                            p_slip = fmaf(-prob_up, prob_down, prob_up + prob_down); // 1-(1-pd)(1-pu)
                            p_upn = prob_up * rcp_fast(prob_up + prob_down);
                        }
                        n0 = ct > 0 ? ct : 0; // below the floor: the stale ct (:478)
                    }
                }
                // mol_amp (:454 / :478), mol_slip (:458), mol_up (:462): ONE copy of the sampler code in the loop --
                // three inlined copies made the sweep ~55 KB of SASS and the warps waited on instruction fetch
                // (ncu: stalled_no_instruction 2.8 per issue, issue-active 60 %)
                int x0 = 0, x1 = 0, x2 = 0;
#pragma unroll 1
                for (int stage = 0; stage < 3; ++stage) {
                    const int dn = stage == 0 ? n0 : (stage == 1 ? (stutter ? x0 : 0) : x1);
                    const float dp = stage == 0 ? eff : (stage == 1 ? p_slip : p_upn);
                    if (!__any_sync(SONICS_FULL, dn > 0))
                        break; // nothing amplified / slipped in any lane: the later stages draw from 0
                    LS_STAT(48 + stage, 1);
                    LS_STAT(52 + stage, __popc(__ballot_sync(SONICS_FULL, dn > 0)));
                    const int x = binomial_ls<FAST>(rng, dn, dp, rcp_tab, a.fast_npq);
                    if (stage == 0)
                        x0 = x;
                    else if (stage == 1)
                        x1 = x;
                    else
                        x2 = x;
                }
                if (in) {
                    cur += x0 - x1;              // products[al] += mol_amp - mol_slip (:466)
                    const int down = x1 - x2;    // mol_down
                    carry = down > 0 ? x2 : 0;   // sic: up-stutter only lands when mol_down > 0 (:473-476)
                    if (b > 0)
                        LS_ST(b - 1, prev + down);
                    if (down > 0 && b - 1 < live_lo)
                        live_lo = b - 1;
                    if (cur != 0 && b > live_hi)
                        live_hi = b;
                    prev = cur;
                    E_up = fmaf(E_up + 1.0f, e1_up, E_up); // expm1((al+1) L) from expm1(al L)
                    E_dn = fmaf(E_dn + 1.0f, e1_dn, E_dn);
                }
            }
            if (bend >= 0)
                LS_ST(bend, prev);
        }

        // ---- lnL (+ readout) of the final pool, all lanes together ----
        if (valid) {
            const GtHeader h = a.t.hdr[g];
            GtView gv;
            gv.W = h.W;
            gv.lo = h.lo;
            gv.A = h.n_alleles;
            gv.D = h.D;
            gv.al = a.t.allele + h.a_off;
            gv.rd = a.t.reads + h.a_off;
            gv.fl = a.t.fl_reads + h.a_off;
            gv.fl_D = h.fl_D;
            int T = 0;
            const double like = pool_lnl(hist, gv, T);
            const int64_t o = (int64_t)g * a.out_stride + a.out_offset + jrel;
            if (!a.replay)
                a.lnL[o] = like;
            if (READOUT) {
                // own RNG stream: the main stream (and therefore lnL) is identical with or without it
                Rng rr;
                rr.init(a.seed, (uint64_t)j, h.uid, STREAM_READOUT);
                double ident;
                float r2;
                pool_readout(hist, gv, T, rr, rcp_tab, ident, r2,
                             a.readout_dump ? a.readout_dump + (size_t)(a.unit_base + u) * Wcap : nullptr);
                finish_store(a, g, o, like, ident, r2);
            }
        }
        __syncwarp();
#ifdef SONICS_LS_STATS
        for (int i = lane; i < 128; i += 32) { // flushed per batch: the 32-bit counters cannot wrap
            atomicAdd(&g_ls_stats[i], (unsigned long long)s_ls_stats[warp][i]);
            s_ls_stats[warp][i] = 0;
        }
        __syncwarp();
#endif
    }
}

} // namespace sonics
