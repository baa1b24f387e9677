// K1: the PCR-stutter simulation kernels (one lane = one simulation at a time).
//
// Restates, per simulation, the reference's
//   generate_params   sonics.pyx:33-45
//   one_repeat        sonics.pyx:293-394  (start alleles, noise, lnL, readout, identity, r^2)
//   simulate          sonics.pyx:397-480  (capture step, amplification / stutter cycles)
// and the lnL of pymc_extracted.f:63-109 (lnl.cuh).
//
// K1a  k_pcr   priors -> start pool -> PCR cycles (+ capture) -> final pool written to a scratch
// K1b  k_lnl   final pool -> sentinel / lnL (fp64 Lanczos) [-> sequencing readout, identity, r^2]
//
// K1a.  The allele-length histogram of a simulation lives in shared memory as a column
// hist[bin * 32 + lane] (bank = lane: conflict free); one cycle is one ascending sweep over the
// lane's live window with the running bin kept in registers:
//   raw   = hist[b]             start-of-cycle count  -> "nzp" snapshot  (:424)
//   cur   = raw + carry         count at visit time, includes the same-cycle
//                               up-stutter carried in from bin b-1            (:429)
//   amplify / slip / up  -> three chained binomials                           (:454-462)
//   down-stutter is added to bin b-1, which is still in a register            (:472)
//   up-stutter becomes `carry` for bin b+1, gated on mol_down > 0 (sic, :473) (:476)
// so a cycle costs one LDS + one STS per live bin on top of the samplers.
//
// Control flow.  Simulations differ in how many draws they need (second start allele, efficiency,
// stutter rates) and every draw in how many rejection trials / inversion steps it takes; ~15 % of
// the BTRS draws need the expensive Stirling test (a quarter of them have n*p < 100).  Lockstep formulations measured 7 of 32
// lanes active (ncu, profiles/r01_k1_v2_ncu.txt) and a warp-per-batch state machine still had 12.5
// lanes idle waiting for the slowest simulation of the warp.  Therefore:
//   * lanes are PERSISTENT: a lane that finishes a simulation stores its pool and immediately pulls
//     the next work unit (genotype, simulation index) from a global counter -- lanes of one warp
//     work on different simulations, possibly of different genotypes, at different cycles;
//   * each lane runs its simulation as a small state machine and the warp iterates over "passes";
//     in one pass a lane performs ADVANCE when its draw has finished and then at most one of
//       ADVANCE  consume the finished draw, walk the sweep to the next draw, set it up
//       INV      set-up + the whole sequential-search inversion      (n*min(p,q) < 20) -- run when
//                >= 6 lanes wait for it
//       BTRS     one transformed-rejection trial incl. squeeze        (otherwise)
//       POIS     one PTRS trial / the small-lambda inversion          (capture step)
//       FULL     the Stirling test of an undecided BTRS candidate -- run only when >= 4 lanes wait
//                for it (or nothing else can progress), so the long path executes with several lanes
//       BOUNDARY end of sweep / end of simulation + next work unit / capture set-up (>= 4 lanes);
//   * the lnL (25 fp64 Lanczos evaluations) is not done by the finishing lane -- that would run at
//     1/32 occupancy -- but by K1b, one lane per simulation, fully convergent.
// Nothing in a simulation's stream depends on neighbouring lanes: random words come from the
// simulation's own Philox counters (rng.cuh), which advance only when that simulation consumes a block, so
// results are independent of launch geometry and any simulation can be replayed bit for bit.
// Reference quirks reproduced: the stale `ct` of the below-floor branch (:478),
// the count/index mix in cap_set (:414), fp32 efficiency/capture/hit (:401),
// up ~ U(up_min, down) (:41), sentinel -999999 (:360-361).
#pragma once
#include <stdint.h>

#include "../../include/sonics_b200.h"
#include "finish.cuh"
#include "lnl.cuh"
#include "rng.cuh"
#include "samplers.cuh"

namespace sonics {

// ---- genotype table (device) ------------------------------------------------
struct GtHeader { // 48 bytes
    int32_t a_off;     // first entry in the allele / reads / fl_reads arrays
    int32_t n_alleles; // non-zero input alleles
    int32_t lo;        // repeat count of window bin 0
    int32_t W;         // live-window width (bins)
    int32_t floor_al;  // sonics.pyx:310-314
    int32_t first_idx; // index of argmax(reads) (first occurrence)           :323
    int32_t D;         // genotype_total = sum(reads)
    int32_t noise_add; // int(float32(noise_coef) * sum(start pool)), 0 = off :334-337
    uint32_t uid;      // RNG stream key
    int32_t in_lo_bin; // window bin of the smallest input allele
    double fl_D;       // factln(D), filled by k_table_prep
};

struct TableView {
    const GtHeader *hdr;
    const int32_t *allele;
    const int32_t *reads;
    const double *fl_reads;
};

struct SimParamsDev {
    int32_t n_cycles, capture_cycle, random_start, up_preference;
    int32_t half_copies; // trunc(start_copies / 2.0)
    double down_lo, down_rng, up_lo, up_hi, cap_lo, cap_rng, eff_lo, eff_rng;
};

// One launch covers the work units u in [0, n_units): unit u = (slot, jrel) with
//   slot = (unit_base + u) / sims_per_gt,  jrel = (unit_base + u) % sims_per_gt,
//   genotype g = active ? active[slot] : slot,  simulation index j = sim_begin + jrel
// (replay mode: one unit per slot, j = replay[g].best_sim).
struct SimArgs {
    TableView t;
    SimParamsDev p;
    const int32_t *active; // nullable
    int64_t unit_base;
    int32_t n_units;
    int32_t sims_per_gt;
    int64_t sim_begin;
    uint64_t seed;
    int64_t out_stride, out_offset;
    double *lnL;
    uint16_t *group;
    double *identity;
    float *rsq;
    double *simpar;
    int32_t Wcap;            // histogram rows per lane (shared memory and scratch)
    int32_t *pool;           // scratch: final pools, [(u / 32) * Wcap + bin] * 32 + (u % 32)
    unsigned int *counter;   // scratch: next unit to hand out
    SonicsVerdict *replay;   // replay mode (K4)
    unsigned long long *dbg; // optional diagnostics (sonics_debug_counters)
    // optional draw trace (sonics_debug_trace; k_pcr<true> only): every draw a simulation requests, in
    // request order, with its outcome (the tests replay the reference's loop on these outcomes)
    SonicsTraceEntry *trace; // [unit][trace_cap]
    int32_t *trace_len;      // [unit] entries requested (may exceed trace_cap: the excess is not stored)
    int32_t trace_cap;
    // optional readout dump (sonics_debug_readout): the simulated sequencing readout per window bin,
    // [unit_base + unit][Wcap]
    int32_t *readout_dump;
    // lockstep kernel (sim_lockstep.cuh): the launch's work units in execution order (nullable = identity)
    const uint32_t *order;
    float fast_npq; // approximate sampler for binomial draws with n p q >= this (k_pcr_ls<.., true> only)
};

// descriptors of the readout -> the per-simulation arrays, or (replay mode, K4) the genotype's verdict
__device__ __forceinline__ void finish_store(const SimArgs &a, int g, int64_t o, double like, double ident, float r2)
{
    if (a.replay) {
        a.replay[g].identity = ident;
        a.replay[g].r_squared = (double)r2;
        if (a.replay[g].lnL != like)
            a.replay[g].best_sim = -2; // replay diverged from the recorded simulation: surfaced by the host
    } else {
        if (a.identity)
            a.identity[o] = ident;
        if (a.rsq)
            a.rsq[o] = r2;
    }
}
enum { K1_CNT_PASSES = 0, K1_CNT_MODE0 = 1 /* .. +6: lanes per mode summed over passes */, K1_CNT_ADV_TRIPS = 8,
       K1_CNT_ADV_LANES = 9, K1_CNT_FULL_RUNS = 10, K1_CNT_FULL_LANES = 11, K1_CNT_WARPS = 12 };

// generate_params (sonics.pyx:33-45) and the start alleles (:320-324) of simulation j of a genotype: the first
// three blocks of the simulation's Philox4x32-10 stream.  One definition for every kernel that needs them (the
// persistent and the lockstep simulation kernels, the sort-key kernel), so they agree bit for bit.
struct Priors {
    double down, up, cap, eff;
    int i1, i2; // indices of the two start alleles in the genotype's allele list
};
__device__ __forceinline__ Priors draw_priors(const SimParamsDev &p, uint64_t seed, const GtHeader &h, int64_t j)
{
    Rng prior;
    prior.init(seed, (uint64_t)j, h.uid, STREAM_MAIN);
    const uint4 q0 = prior.block4(), q1 = prior.block4(), q2 = prior.block4();
    Priors r;
    // d, u, c, p in that order
    // low + (high - low) * u, product and sum rounded separately as numpy's uniform() does
    auto uni = [](double lo, double rng, double u) { return __dadd_rn(lo, __dmul_rn(rng, u)); };
    r.down = uni(p.down_lo, p.down_rng, u01_53(q0.x, q0.y));
    r.up = uni(p.up_lo, (p.up_preference ? p.up_hi : r.down) - p.up_lo, u01_53(q0.z, q0.w));
    r.cap = uni(p.cap_lo, p.cap_rng, u01_53(q1.x, q1.y));
    r.eff = uni(p.eff_lo, p.eff_rng, u01_53(q1.z, q1.w));
    const uint32_t A = (uint32_t)h.n_alleles;
    if (p.random_start) {
        r.i1 = (int)__umulhi(q2.x, A);
        r.i2 = (int)__umulhi(q2.y, A);
    } else {
        r.i1 = h.first_idx;
        r.i2 = (int)__umulhi(q2.x, A);
    }
    return r;
}

// work unit u of a launch -> (slot, jrel): 32-bit division when it fits (the 64-bit one is a ~100-instruction call)
__device__ __forceinline__ void unit_decode(const SimArgs &a, int u, int &slot, int &jrel)
{
    const int64_t gu = a.unit_base + u;
    slot = gu < 0x7fffffffLL ? (int)((uint32_t)gu / (uint32_t)a.sims_per_gt) : (int)(gu / a.sims_per_gt);
    jrel = (int)(gu - (int64_t)slot * a.sims_per_gt);
}

// fl_reads[i] = factln(reads[i]); hdr[g].fl_D = factln(D)
__global__ void k_table_prep(GtHeader *hdr, const int32_t *reads, double *fl_reads, int n_gt, int n_alleles)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_alleles)
        fl_reads[i] = factln_dev(reads[i]);
    if (i < n_gt)
        hdr[i].fl_D = factln_dev(hdr[i].D);
}

// pymc_extracted.mvhyperg on arbitrary (x, color) rows: one thread per row, the
// Fortran loop verbatim (pymc_extracted.f:88-108).
__global__ void k_mvhyperg(const int64_t *x, const int64_t *color, int k, int n_pools, double *out)
{
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_pools)
        return;
    const int64_t *xr = x + (size_t)r * k;
    const int64_t *cr = color + (size_t)r * k;
    int d = 0, total = 0;
    double like = 0.0;
    for (int i = 0; i < k; ++i) {
        const int xi = (int)xr[i], ci = (int)cr[i];
        like = __dadd_rn(like, factln_dev(ci));
        like = __dsub_rn(like, factln_dev(xi));
        like = __dsub_rn(like, factln_dev(ci - xi));
        if (ci < 0 || xi < 0) {
            out[r] = SONICS_NEG_INFINITY;
            return;
        }
        d += xi;
        total += ci;
    }
    if (total <= 0) {
        out[r] = SONICS_NEG_INFINITY;
        return;
    }
    like = __dsub_rn(like, __dsub_rn(__dsub_rn(factln_dev(total), factln_dev(d)), factln_dev(total - d)));
    out[r] = like;
}

// ---- test hooks ----
__global__ void k_debug_philox(uint64_t seed, uint64_t sim, uint32_t uid, uint32_t stream_id, int n_blocks,
                               uint32_t *out)
{
    Rng r;
    r.init(seed, sim, uid, stream_id);
    for (int i = 0; i < n_blocks; ++i) {
        const uint4 v = r.block4();
        out[4 * i + 0] = v.x;
        out[4 * i + 1] = v.y;
        out[4 * i + 2] = v.z;
        out[4 * i + 3] = v.w;
    }
}

__global__ void k_debug_pairs(uint64_t seed, uint64_t sim, uint32_t uid, int n_pairs, uint32_t *out)
{
    PairRng r;
    r.init(seed, sim, uid);
    for (int i = 0; i < n_pairs; ++i) {
        const uint2 v = r.pair();
        out[2 * i + 0] = v.x;
        out[2 * i + 1] = v.y;
    }
}

__global__ void k_debug_sample(int kind, int n, double p, int64_t count, uint64_t seed, int32_t *out)
{
    __shared__ float rcp_tab[SONICS_RCP_TABLE];
    for (int i = threadIdx.x; i < SONICS_RCP_TABLE; i += blockDim.x)
        rcp_tab[i] = i ? 1.0f / (float)i : 0.0f;
    __syncthreads();
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= count)
        return;
    Rng r;
    r.init(seed, (uint64_t)t, 0x5eedu, STREAM_MAIN);
    if (kind == 0) {
        const uint4 b = r.block4();
        out[t] = binomial(r, n, (float)p, b.x, b.y, rcp_tab);
    } else if (kind == 2) { // the approximate sampler, as binomial_ls<true> calls it
        const uint4 b = r.block4();
        const float pf = (float)p;
        const int k = binomial_normal(n, pf > 0.5f ? 1.0f - pf : pf, b.x, b.y);
        out[t] = pf > 0.5f ? n - k : k;
    } else {
        out[t] = poisson(r, (float)p);
    }
}

#define SONICS_FULL 0xffffffffu

// Shared-memory layout of k_pcr: [rcp table: SONICS_RCP_TABLE floats][warp 0 hist][warp 1 hist]...
__host__ __device__ inline size_t warp_smem_bytes(int Wcap) { return (size_t)Wcap * 32 * 4; }

__device__ __forceinline__ size_t pool_index(int u, int Wcap, int bin)
{
    return ((size_t)(u >> 5) * Wcap + bin) * 32 + (u & 31);
}

enum { M_ADV = 0, M_BTRS = 1, M_INV0 = 2, M_BOUND = 3, M_FULL = 4, M_POIS = 5, M_EXIT = 6 };
enum { PC_NEWSIM = 0, PC_RESULT = 1, PC_CAPRESULT = 2, PC_SCAN = 3 };
// deferral thresholds: a minority phase runs when this many lanes wait for it (or nothing else can progress)
#ifndef SONICS_THR_BOUND
#define SONICS_THR_BOUND 4
#endif
#ifndef SONICS_THR_INV
#define SONICS_THR_INV 6
#endif
#ifndef SONICS_THR_FULL
#define SONICS_THR_FULL 4
#endif
#ifndef SONICS_PCR_MIN_BLOCKS
#define SONICS_PCR_MIN_BLOCKS 7 // x 128 threads x 72 registers; 8 x 64 spills 80 B per thread and is 1.6 % slower
#endif

// ---------------------------------------------------------------------------------------------
// K1a
// ---------------------------------------------------------------------------------------------
// TRACE: the instrumented instantiation (control-flow counters of sonics_debug_counters and the draw trace
// of sonics_debug_trace); the production instantiation carries neither (the counters alone cost 5 %).
template <bool TRACE>
__global__ void __launch_bounds__(128, SONICS_PCR_MIN_BLOCKS) k_pcr(const SimArgs a)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int Wcap = a.Wcap;
    float *rcp_tab = reinterpret_cast<float *>(smem_raw);
    for (int i = threadIdx.x; i < SONICS_RCP_TABLE; i += blockDim.x)
        rcp_tab[i] = i ? 1.0f / (float)i : 0.0f;
    __syncthreads();
    int *hist = reinterpret_cast<int *>(smem_raw + SONICS_RCP_TABLE * 4 + (size_t)warp * warp_smem_bytes(Wcap)) + lane;

    const int n_cycles = a.p.n_cycles, cap_cycle = a.p.capture_cycle;

    // ---- per-lane simulation state ----
    int mode = M_BOUND, pc = PC_NEWSIM;
    int u = -1; // work unit in flight
    int W = 0, lo = 0, floor_bin = 0;
    PairRng rng;
    rng.init(a.seed, 0, 0);
    float eff = 0.f, capture = 0.f, l_up = 0.f, l_dn = 0.f, e1_up = 0.f, e1_dn = 0.f;
    int live_lo = 0, live_hi = -1;
    int cycle = 1;
    int ct = 0; // function-scope `ct` of the reference; goes stale below the floor (:478)
    int b = 1, bend = 0, prev = 0, cur = 0, carry = 0;
    int stage = 0, n_stage = 1;
    float p_slip = 0.0f, p_upn = 0.0f, E_up = 0.0f, E_dn = 0.0f;
    int x = 0;         // result of the draw that just finished
    int dn = 0;        // n of the draw in flight
    bool flip = false; // draw in flight is on 1 - p
    bool sweeping = false, in_capture = false, cap_done = false;
    int cap_b = 0, cap_hi = 0, cap_n = 1, cap_v = 0;
    float cap_set = 0.0f, floor_cap = 0.0f;
    DrawState ds;
    int tr_n = 0; // TRACE: draws requested by the simulation in flight

    unsigned long long c_pass = 0, c_adv_trips = 0, c_adv_lanes = 0, c_full_runs = 0, c_full_lanes = 0;
    unsigned long long c_mode[7] = {0, 0, 0, 0, 0, 0, 0};

    while (true) {
        // ================= BOUNDARY (deferred) =================
        // end of a sweep -> next cycle; end of a simulation -> store the pool, pull the next work unit
        // and initialise it; capture bookkeeping; begin the sweep of the cycle.
        {
            const unsigned m_bound = __ballot_sync(SONICS_FULL, mode == M_BOUND);
            const unsigned m_busy = __ballot_sync(SONICS_FULL, mode == M_ADV || mode == M_BTRS || mode == M_POIS);
            if (__popc(m_bound) >= SONICS_THR_BOUND || (m_bound != 0u && m_busy == 0u)) {
                if (mode == M_BOUND) {
                    if (pc != PC_NEWSIM) {
                        if (sweeping) { // store the last bin of the sweep, move to the next cycle
                            hist[bend * 32] = prev;
                            sweeping = false;
                            ++cycle;
                        }
                        if (cycle > n_cycles) { // simulation finished: hand the pool to K1b
                            for (int c = 0; c < W; ++c)
                                a.pool[pool_index(u, Wcap, c)] = hist[c * 32];
                            if (TRACE && a.trace)
                                a.trace_len[u] = tr_n;
                            pc = PC_NEWSIM;
                        }
                    }
                    if (pc == PC_NEWSIM) {
                        const unsigned int t = atomicAdd(a.counter, 1u);
                        if (t >= (unsigned int)a.n_units) {
                            mode = M_EXIT;
                        } else {
                            u = (int)t;
                            tr_n = 0;
                            int slot, jrel;
                            unit_decode(a, u, slot, jrel);
                            const int g = a.active ? a.active[slot] : slot;
                            const GtHeader h = a.t.hdr[g];
                            int64_t j = a.sim_begin + jrel;
                            if (a.replay) {
                                j = a.replay[g].best_sim;
                                if (j < 0)
                                    j = 0; // nothing to replay: simulate index 0, K1b ignores the result
                            }
                            W = h.W;
                            lo = h.lo;
                            floor_bin = h.floor_al - lo;
                            const int A = h.n_alleles;
                            const Priors pr = draw_priors(a.p, a.seed, h, j);
                            rng.init(a.seed, (uint64_t)j, h.uid); // the draws of the PCR cycles
                            const double par_down = pr.down, par_up = pr.up, par_cap = pr.cap, par_eff = pr.eff;
                            eff = (float)par_eff; // cdef float efficiency, capture (:401)
                            capture = (float)par_cap;
                            l_up = log1pf(-(float)par_up);   // ln(1 - up)
                            l_dn = log1pf(-(float)par_down); // ln(1 - down)
                            e1_up = expm1f(l_up);            // per-bin step of expm1(al * L)
                            e1_dn = expm1f(l_dn);
                            const int i1 = pr.i1, i2 = pr.i2; // start alleles (:320-332)
                            const int b1 = a.t.allele[h.a_off + i1] - lo;
                            const int b2 = a.t.allele[h.a_off + i2] - lo;
                            // initial pool (:326-337)
                            for (int c = 0; c < W; ++c)
                                hist[c * 32] = 0;
                            hist[b1 * 32] += a.p.half_copies;
                            hist[b2 * 32] += a.p.half_copies;
                            live_lo = min(b1, b2);
                            live_hi = max(b1, b2);
                            if (h.noise_add > 0) {
                                for (int i = 0; i < A; ++i)
                                    hist[(a.t.allele[h.a_off + i] - lo) * 32] += h.noise_add;
                                live_lo = h.in_lo_bin;
                                live_hi = a.t.allele[h.a_off + A - 1] - lo;
                            }
                            if (!a.replay) {
                                const int64_t o = (int64_t)g * a.out_stride + a.out_offset + jrel;
                                a.group[o] = (uint16_t)(min(i1, i2) * A + max(i1, i2));
                                if (a.simpar) {
                                    a.simpar[o * 4 + 0] = par_down;
                                    a.simpar[o * 4 + 1] = par_up;
                                    a.simpar[o * 4 + 2] = par_cap;
                                    a.simpar[o * 4 + 3] = par_eff;
                                }
                            }
                            cycle = 1;
                            ct = 0;
                            carry = 0;
                            cap_done = false;
                            in_capture = false;
                            sweeping = false;
                            pc = PC_SCAN;
                        }
                    }
                    if (mode != M_EXIT && cycle > n_cycles) {
                        // n_cycles == 0 (range(1, 1) runs no cycle, sonics.pyx:410): the start pool is the result
                        for (int c = 0; c < W; ++c)
                            a.pool[pool_index(u, Wcap, c)] = hist[c * 32];
                        if (TRACE && a.trace)
                            a.trace_len[u] = tr_n;
                        pc = PC_NEWSIM; // stays in M_BOUND: the next BOUNDARY pass pulls the next unit
                    } else if (mode != M_EXIT) {
                        bool to_capture = false;
                        if (cycle == cap_cycle && !cap_done) {
                            // capture step (:412-423): max count, first non-zero allele, non-zero count
                            cap_done = true;
                            int maxp = 0, nnz = 0, minidx = -1;
                            for (int c = live_lo; c <= live_hi; ++c) {
                                const int v = hist[c * 32];
                                maxp = max(maxp, v);
                                if (v != 0) {
                                    ++nnz;
                                    if (minidx < 0)
                                        minidx = lo + c;
                                }
                            }
                            if (nnz > 0) {
                                cap_set = (float)((double)capture / (double)(maxp - minidx));
                                floor_cap = (float)(1.0 - ((double)__fmul_rn(cap_set, (float)nnz)) / 2.0);
                                cap_n = 1;
                                cap_b = live_lo;
                                cap_hi = live_hi;
                                in_capture = true;
                                to_capture = true;
                            }
                        }
                        if (!to_capture) {
                            // begin the amplification sweep of this cycle (:424-479)
                            sweeping = true;
                            b = live_lo;
                            bend = min(live_hi + 1, W - 1);
                            carry = 0;
                            prev = b > 0 ? hist[(b - 1) * 32] : 0;
                            const float al = (float)(lo + b);
                            E_up = expm1f(__fmul_rn(al, l_up));
                            E_dn = expm1f(__fmul_rn(al, l_dn));
                        }
                        mode = M_ADV;
                        pc = PC_SCAN;
                    }
                }
            }
        }

        // ================= ADVANCE =================
        if (TRACE && a.dbg)
            c_adv_lanes += __popc(__ballot_sync(SONICS_FULL, mode == M_ADV));
        const unsigned m_adv = __ballot_sync(SONICS_FULL, mode == M_ADV);
        if (mode == M_ADV) {
            int req = 0; // 1: binomial(dn, dp)   2: poisson(lam)
            float dp = 0.0f, lam = 0.0f;
            if (TRACE && a.trace && (pc == PC_RESULT || pc == PC_CAPRESULT) && tr_n <= a.trace_cap)
                a.trace[(size_t)u * a.trace_cap + tr_n - 1].x = x;
            // ---- (1) consume the finished draw ----
            // Written with selects instead of a branch per stage: the lanes of a pass are spread over
            // the three stages, and a 3-way branch would run each third separately.
            if (pc == PC_RESULT) {
                // mol_amp joins the bin as soon as it is known (stage 0), mol_slip leaves it (stage 1) and waits
                // in `carry` -- zero for the duration of a visit -- for the third draw: no per-draw state beyond
                // what the sweep carries anyway (products[al] += mol_amp - mol_slip, :466)
                cur += stage == 0 ? x : (stage == 1 ? -x : 0);
                carry = stage == 1 ? x : carry;
                // the bin is finished after the third draw, after a draw that returned 0 (nothing
                // amplified / nothing slipped), or after the single draw of a below-floor bin (:477-479)
                const bool bin_done = stage == 2 || x == 0 || n_stage == 1;
                if (bin_done) {
                    const int up_e = stage == 2 ? x : 0;
                    const int down = carry - up_e; // mol_down = mol_slip - mol_up
                    carry = down > 0 ? up_e : 0;   // sic: up-stutter only lands when mol_down > 0 (:473-476)
                    if (b > 0)
                        hist[(b - 1) * 32] = prev + down;
                    if (down > 0 && b - 1 < live_lo)
                        live_lo = b - 1;
                    if (cur != 0 && b > live_hi)
                        live_hi = b;
                    prev = cur;
                    ++b;
                    E_up = fmaf(E_up + 1.0f, e1_up, E_up); // expm1((al+1) L) from expm1(al L)
                    E_dn = fmaf(E_dn + 1.0f, e1_dn, E_dn);
                    pc = PC_SCAN;
                } else {
                    ++stage;
                    dn = x;
                    dp = stage == 1 ? p_slip : p_upn;
                    req = 1;
                }
            } else if (pc == PC_CAPRESULT) {
                hist[cap_b * 32] = x < cap_v ? x : cap_v; // ct_up = min(poisson(hit), ct) (:421)
                ++cap_n;
                ++cap_b;
                pc = PC_SCAN;
            }
            // ---- (2) walk to the next draw ----
            while (req == 0 && mode == M_ADV) {
                if (TRACE && a.dbg && lane == (__ffs(__activemask()) - 1))
                    ++c_adv_trips;
                if (in_capture) {
                    // capture step (:412-423): one Poisson draw per non-zero bin
                    if (cap_b > cap_hi) {
                        in_capture = false;
                        mode = M_BOUND; // begin the sweep of this same cycle
                        break;
                    }
                    cap_v = hist[cap_b * 32];
                    if (cap_v != 0) {
                        ct = cap_v;
                        lam = __fmul_rn(__fadd_rn(floor_cap, __fmul_rn(cap_set, (float)cap_n)), (float)cap_v);
                        pc = PC_CAPRESULT;
                        if (lam > 0.0f) {
                            req = 2;
                        } else { // numpy raises for lam < 0; unreachable with sane ranges: treat as 0
                            hist[cap_b * 32] = 0;
                            ++cap_n;
                            ++cap_b;
                            pc = PC_SCAN;
                        }
                    } else {
                        ++cap_b;
                    }
                    continue;
                }
                if (b > bend) {
                    mode = M_BOUND;
                    break;
                }
                // visit bin b
                const int raw = hist[b * 32];
                cur = raw + carry;
                carry = 0;
                bool skip = raw == 0; // empty at cycle start: not amplified this cycle (:424)
                if (!skip) {
                    const bool stutter = b >= floor_bin;
                    if (stutter) {
                        ct = cur;
                        const float prob_up = -E_up, prob_down = -E_dn;
                        p_slip = fmaf(-prob_up, prob_down, prob_up + prob_down); // 1-(1-pd)(1-pu)
                        p_upn = prob_up * rcp_fast(prob_up + prob_down);
                    }
                    n_stage = stutter ? 3 : 1;
                    stage = 0;
                    if (ct > 0) {
                        dn = ct;
                        dp = eff;
                        req = 1;
                        pc = PC_RESULT;
                    } else {
                        skip = true; // stale ct == 0 below the floor: Binomial(0, eff) = 0
                    }
                }
                if (skip) {
                    if (b > 0)
                        hist[(b - 1) * 32] = prev;
                    if (cur != 0 && b > live_hi)
                        live_hi = b;
                    prev = cur;
                    ++b;
                    E_up = fmaf(E_up + 1.0f, e1_up, E_up);
                    E_dn = fmaf(E_dn + 1.0f, e1_dn, E_dn);
                }
            }
            // ---- (3) classify the draw ----
            __syncwarp(m_adv); // lanes leave the walk loop at different trips: classify / set up together (+8 %)
            if (TRACE && a.trace && req != 0) {
                if (tr_n < a.trace_cap) {
                    SonicsTraceEntry e;
                    e.kind = req;
                    e.n = req == 1 ? dn : cap_v;
                    e.p = req == 1 ? dp : lam;
                    e.x = -1;
                    a.trace[(size_t)u * a.trace_cap + tr_n] = e;
                }
                ++tr_n;
            }
            if (req == 1) {
                if (!(dp > 0.0f)) { // degenerate priors only: consumed on the next pass without a draw
                    x = 0;
                } else if (dp >= 1.0f) {
                    x = dn;
                } else {
                    flip = dp > 0.5f;
                    const float pp = flip ? 1.0f - dp : dp;
                    if ((float)dn * pp < SONICS_INV_BELOW) {
                        ds.p = pp; // set up when the INV phase runs
                        mode = M_INV0;
                    } else {
                        btrs_setup(ds, dn, pp);
                        mode = M_BTRS;
                    }
                }
            } else if (req == 2) {
                if (lam >= 10.0f)
                    pois_setup(ds, lam);
                else
                    ds.fn = lam;
                mode = M_POIS;
            }
        }
        if (__all_sync(SONICS_FULL, mode == M_EXIT))
            break;
        if (TRACE && a.dbg) {
            ++c_pass;
#pragma unroll
            for (int mm = 0; mm < 7; ++mm)
                c_mode[mm] += __popc(__ballot_sync(SONICS_FULL, mode == mm));
        }
        const unsigned m_busy = __ballot_sync(SONICS_FULL, mode == M_ADV || mode == M_BTRS || mode == M_POIS);

        // ================= INV (deferred) + one rejection step: shared random-word fetch =================
        // Small-mean binomials are drawn by inversion when enough lanes wait for it (set-up and the whole
        // sequential search in that pass: a pass costs a waiting lane more than the search's tail).  Lanes in a rejection sampler run ONE trial per pass (an immediate
        // retry would execute with ~1 lane).  Every lane that consumes randomness in this pass takes
        // the next 64-bit block of ITS OWN simulation's stream -- the block counter advances only
        // when the simulation consumes, never with the warp's pass count -- through one call site.
        {
            const unsigned m_inv = __ballot_sync(SONICS_FULL, mode == M_INV0);
            const bool run_inv = __popc(m_inv) >= SONICS_THR_INV || (m_inv != 0u && m_busy == 0u);
            uint2 words = make_uint2(0u, 0u);
            if (mode == M_BTRS || mode == M_POIS || (run_inv && mode == M_INV0))
                words = rng.pair();
            if (run_inv && mode == M_INV0) {
                inv_setup(ds, dn, ds.p);
                inv_start(ds, words.x);
                inv_run(ds, rcp_tab);
                x = flip ? dn - ds.ix : ds.ix;
                mode = M_ADV;
            } else if (mode == M_BTRS) {
                const int r = btrs_trial(ds, words.x, words.y);
                if (r == 1) {
                    const int k = (int)ds.fk;
                    x = flip ? dn - k : k;
                    mode = M_ADV;
                } else if (r == 2) {
                    mode = M_FULL;
                }
            } else if (mode == M_POIS) {
                if (ds.fn < 10.0f) {
                    x = poisson_small(ds.fn, words.x);
                    mode = M_ADV;
                } else if (pois_trial(ds, words.x, words.y)) {
                    x = ds.ix;
                    mode = M_ADV;
                }
            }
        }

        // ================= deferred Stirling test =================
        {
            const unsigned m_full = __ballot_sync(SONICS_FULL, mode == M_FULL);
            if (__popc(m_full) >= SONICS_THR_FULL || (m_full != 0u && m_busy == 0u)) {
                if (TRACE && a.dbg) {
                    ++c_full_runs;
                    c_full_lanes += __popc(m_full);
                }
                if (mode == M_FULL) {
                    if (btrs_full(ds)) {
                        const int k = (int)ds.fk;
                        x = flip ? dn - k : k;
                        mode = M_ADV;
                    } else {
                        mode = M_BTRS;
                    }
                }
            }
        }
    }

    if (TRACE && a.dbg) {
        for (int o = 16; o > 0; o >>= 1)
            c_adv_trips += __shfl_down_sync(SONICS_FULL, c_adv_trips, o);
        if (lane == 0) {
            atomicAdd(&a.dbg[K1_CNT_PASSES], c_pass);
            for (int mm = 0; mm < 7; ++mm)
                atomicAdd(&a.dbg[K1_CNT_MODE0 + mm], c_mode[mm]);
            atomicAdd(&a.dbg[K1_CNT_ADV_TRIPS], c_adv_trips);
            atomicAdd(&a.dbg[K1_CNT_ADV_LANES], c_adv_lanes);
            atomicAdd(&a.dbg[K1_CNT_FULL_RUNS], c_full_runs);
            atomicAdd(&a.dbg[K1_CNT_FULL_LANES], c_full_lanes);
            atomicAdd(&a.dbg[K1_CNT_WARPS], 1ull);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// K1b: one lane per simulation -- lnL of the final pool (:360-363, pymc_extracted.f:88-108) and,
// with READOUT, the sequencing readout and its descriptors (:341-357, :365-384, rsq :273-291).
// Lanes of a warp read consecutive columns of the pool scratch (coalesced) and, for units of the
// same genotype, the same table entries (broadcast).
// ---------------------------------------------------------------------------------------------
template <bool READOUT>
__global__ void __launch_bounds__(128) k_lnl(const SimArgs a)
{
    __shared__ float rcp_tab[READOUT ? SONICS_RCP_TABLE : 1];
    if (READOUT) {
        for (int i = threadIdx.x; i < SONICS_RCP_TABLE; i += blockDim.x)
            rcp_tab[i] = i ? 1.0f / (float)i : 0.0f;
        __syncthreads();
    }
    const int Wcap = a.Wcap;
    for (int u = blockIdx.x * blockDim.x + threadIdx.x; u < a.n_units; u += gridDim.x * blockDim.x) {
        const int64_t gu = a.unit_base + u;
        const int slot = (int)(gu / a.sims_per_gt);
        const int jrel = (int)(gu - (int64_t)slot * a.sims_per_gt);
        const int g = a.active ? a.active[slot] : slot;
        const GtHeader h = a.t.hdr[g];
        int64_t j = a.sim_begin + jrel;
        if (a.replay) {
            j = a.replay[g].best_sim;
            if (j < 0)
                continue;
        }
        int *col = a.pool + pool_index(u, Wcap, 0); // bin c of this simulation's pool is col[c * 32]
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
        const double like = pool_lnl(col, gv, T);
        const int64_t o = (int64_t)g * a.out_stride + a.out_offset + jrel;
        if (!a.replay)
            a.lnL[o] = like;

        if (READOUT) {
            // own RNG stream: the main stream (and therefore lnL) is identical with or without it
            Rng rr;
            rr.init(a.seed, (uint64_t)j, h.uid, STREAM_READOUT);
            double ident;
            float r2;
            pool_readout(col, gv, T, rr, rcp_tab, ident, r2,
                         a.readout_dump ? a.readout_dump + (size_t)gu * Wcap : nullptr);
            finish_store(a, g, o, like, ident, r2);
        }
    }
}

} // namespace sonics
