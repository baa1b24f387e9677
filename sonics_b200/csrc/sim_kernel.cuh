// K1: the fused PCR-stutter simulation kernel (one lane = one simulation).
//
// Restates, per simulation, the reference's
//   generate_params   sonics.pyx:33-45
//   one_repeat        sonics.pyx:293-394  (start alleles, noise, lnL, readout, identity, r^2)
//   simulate          sonics.pyx:397-480  (capture step, amplification / stutter cycles)
// and the lnL of pymc_extracted.f:63-109 (lnl.cuh).
//
// Mapping onto the GPU.  All 32 lanes of a warp simulate the SAME genotype (same reads, same
// window), so the genotype description is staged once per warp in shared memory.  The
// allele-length histogram of each simulation lives in shared memory as a column
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
// Control flow (r01 v3).  The draws of different simulations need different numbers of rejection
// trials / inversion steps, and ~2 % of the binomial trials need the expensive Stirling test.  In
// a lockstep formulation every such event stalls the other 31 lanes: ncu measured 7 of 32 lanes
// active on average.  Here each lane runs its simulation as a small STATE MACHINE and the warp
// iterates over "passes"; in one pass a lane performs at most one of
//   ADVANCE   finish the previous draw, walk the sweep to the next non-trivial draw, set it up
//   INV       up to 8 steps of sequential-search inversion        (n*min(p,q) < 10)
//   BTRS      one transformed-rejection trial incl. the squeeze   (otherwise)
//   POIS      one PTRS trial (capture step)
//   FULL      the Stirling test of an undecided BTRS candidate -- run only when >= 4 lanes wait for
//             it (or nothing else can progress), so the long path executes with several lanes
// Lanes drift apart by design; nothing in a lane's stream depends on its neighbours (random words
// come from the lane's own Philox counter, a fresh 128-bit block every second pass), so results are
// independent of launch geometry and a single simulation can be replayed bit for bit.
// Reference quirks reproduced: the stale `ct` of the below-floor branch (:478),
// the count/index mix in cap_set (:414), fp32 efficiency/capture/hit (:401),
// up ~ U(up_min, down) (:41), sentinel -999999 (:360-361).
#pragma once
#include <stdint.h>

#include "../../include/sonics_b200.h"
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

struct SimArgs {
    TableView t;
    SimParamsDev p;
    const int32_t *active; // nullable
    int32_t n_active;
    int32_t warps_per_gt;
    int64_t sim_begin, sim_count;
    uint64_t seed;
    int64_t out_stride, out_offset;
    double *lnL;
    uint16_t *group;
    double *identity;
    float *rsq;
    double *simpar;
    int32_t Wcap; // histogram rows per warp in shared memory
    // replay mode (K4): one simulation per genotype, index taken from verdict[g].best_sim
    SonicsVerdict *replay;
    // optional diagnostics (sonics_debug_counters): 16 x u64, see K1_CNT_* below
    unsigned long long *dbg;
};
enum { K1_CNT_PASSES = 0, K1_CNT_MODE0 = 1 /* .. +6: lanes per mode summed over passes */, K1_CNT_ADV_TRIPS = 8,
       K1_CNT_ADV_LANES = 9, K1_CNT_FULL_RUNS = 10, K1_CNT_FULL_LANES = 11, K1_CNT_ITEMS = 12 };

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
    } else {
        out[t] = poisson(r, p);
    }
}

#define SONICS_FULL 0xffffffffu

// Shared-memory layout of k_simulate: [rcp table: SONICS_RCP_TABLE floats][warp 0 region][warp 1 region]...
// Shared-memory bytes one warp needs for a window of Wcap bins.
__host__ __device__ inline size_t warp_smem_bytes(int Wcap)
{
    // hist: Wcap*32 int32 | reads: Wcap int32 (padded to 8 B) | fl_reads: Wcap double
    return (size_t)Wcap * 32 * 4 + (size_t)((Wcap + 1) & ~1) * 4 + (size_t)Wcap * 8;
}

enum { M_ADV = 0, M_BTRS = 1, M_INV0 = 2, M_INV = 3, M_FULL = 4, M_POIS = 5, M_DONE = 6 };
enum { PC_CYCLE = 0, PC_CAPSCAN = 1, PC_CAPRESULT = 2, PC_BINSCAN = 3, PC_RESULT = 4, PC_FINISH = 5 };

template <bool READOUT>
__global__ void __launch_bounds__(128) k_simulate(const SimArgs a)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int warps_per_block = blockDim.x >> 5;
    const int Wcap = a.Wcap;
    float *rcp_tab = reinterpret_cast<float *>(smem_raw);
    for (int i = threadIdx.x; i < SONICS_RCP_TABLE; i += blockDim.x)
        rcp_tab[i] = i ? 1.0f / (float)i : 0.0f;
    __syncthreads();
    unsigned char *wbase = smem_raw + SONICS_RCP_TABLE * 4 + (size_t)warp * warp_smem_bytes(Wcap);
    int *hist = reinterpret_cast<int *>(wbase) + lane; // column: hist[b * 32]
    int *s_reads = reinterpret_cast<int *>(wbase) + Wcap * 32;
    double *s_fl = reinterpret_cast<double *>(wbase + (size_t)Wcap * 32 * 4 + (size_t)((Wcap + 1) & ~1) * 4);

    const int64_t n_items = (int64_t)a.n_active * a.warps_per_gt;
    for (int64_t item = (int64_t)blockIdx.x * warps_per_block + warp; item < n_items;
         item += (int64_t)gridDim.x * warps_per_block) {
        const int slot = (int)(item / a.warps_per_gt);
        const int wi = (int)(item - (int64_t)slot * a.warps_per_gt);
        const int g = a.active ? a.active[slot] : slot;
        const GtHeader h = a.t.hdr[g];
        const int W = h.W, lo = h.lo, A = h.n_alleles, D = h.D;
        int64_t j = a.sim_begin + (int64_t)wi * 32 + lane;
        bool valid = j < a.sim_begin + a.sim_count;
        if (a.replay) {
            j = a.replay[g].best_sim;
            valid = lane == 0 && j >= 0;
            if (j < 0)
                j = 0;
        }

        // ---- stage the genotype (warp-uniform data) ----
        __syncwarp();
        for (int b = lane; b < W; b += 32) {
            s_reads[b] = 0;
            s_fl[b] = 0.0;
        }
        __syncwarp();
        for (int i = lane; i < A; i += 32) {
            const int bi = a.t.allele[h.a_off + i] - lo;
            s_reads[bi] = a.t.reads[h.a_off + i];
            s_fl[bi] = a.t.fl_reads[h.a_off + i];
        }
        __syncwarp();
        const int in_lo = h.in_lo_bin;
        const int in_hi = a.t.allele[h.a_off + A - 1] - lo;

        // ---- generate_params (sonics.pyx:33-45): d, u, c, p in that order ----
        Rng rng;
        rng.init(a.seed, (uint64_t)j, h.uid, STREAM_MAIN);
        const uint4 q0 = rng.block4(), q1 = rng.block4(), q2 = rng.block4();
        const double par_down = a.p.down_lo + a.p.down_rng * u01_53(q0.x, q0.y);
        const double par_up = a.p.up_preference ? a.p.up_lo + (a.p.up_hi - a.p.up_lo) * u01_53(q0.z, q0.w)
                                                : a.p.up_lo + (par_down - a.p.up_lo) * u01_53(q0.z, q0.w);
        const double par_cap = a.p.cap_lo + a.p.cap_rng * u01_53(q1.x, q1.y);
        const double par_eff = a.p.eff_lo + a.p.eff_rng * u01_53(q1.z, q1.w);
        const float eff = (float)par_eff; // cdef float efficiency, capture (:401)
        const float capture = (float)par_cap;
        const float l_up = log1pf(-(float)par_up);   // ln(1 - up)
        const float l_dn = log1pf(-(float)par_down); // ln(1 - down)
        const float e1_up = expm1f(l_up), e1_dn = expm1f(l_dn); // per-bin step of expm1(al * L)

        // ---- start alleles (:320-332) ----
        int i1, i2;
        if (a.p.random_start) {
            i1 = (int)__umulhi(q2.x, (uint32_t)A);
            i2 = (int)__umulhi(q2.y, (uint32_t)A);
        } else {
            i1 = h.first_idx;
            i2 = (int)__umulhi(q2.x, (uint32_t)A);
        }
        const int b1 = a.t.allele[h.a_off + i1] - lo;
        const int b2 = a.t.allele[h.a_off + i2] - lo;

        // ---- initial pool (:326-337) ----
        for (int b = 0; b < W; ++b)
            hist[b * 32] = 0;
        int live_lo, live_hi;
        if (valid) {
            hist[b1 * 32] += a.p.half_copies;
            hist[b2 * 32] += a.p.half_copies;
            live_lo = min(b1, b2);
            live_hi = max(b1, b2);
            if (h.noise_add > 0) {
                for (int b = in_lo; b <= in_hi; ++b)
                    if (s_reads[b] != 0)
                        hist[b * 32] += h.noise_add;
                live_lo = in_lo;
                live_hi = in_hi;
            }
        } else { // idle lane of a partially filled warp
            live_lo = W;
            live_hi = -1;
        }

        // ---- simulate (:397-480) as a per-lane state machine ----
        const int floor_bin = h.floor_al - lo;
        const int n_cycles = a.p.n_cycles, cap_cycle = a.p.capture_cycle;
        int mode = valid ? M_ADV : M_DONE;
        int pc = PC_CYCLE;
        int cycle = 1;
        int ct = 0; // function-scope `ct` of the reference; goes stale below the floor (:478)
        int b = 0, bend = 0, prev = 0, cur = 0, carry = 0, down = 0;
        int stage = 0, n_stage = 1, amp = 0, slip = 0;
        float p_slip = 0.0f, p_upn = 0.0f, E_up = 0.0f, E_dn = 0.0f;
        int x = 0;          // result of the draw that just finished
        int dn = 0;         // n of the draw in flight
        bool flip = false;  // draw in flight is on 1 - p
        bool cap_done = false;
        int cap_b = 0, cap_hi = 0, cap_n = 1, cap_v = 0;
        float cap_set = 0.0f, floor_cap = 0.0f;
        BtrsState bs;
        InvState iv;
        PoisState ps;

        unsigned long long c_pass = 0, c_adv_trips = 0, c_adv_lanes = 0, c_full_runs = 0, c_full_lanes = 0;
        unsigned long long c_mode[7] = {0, 0, 0, 0, 0, 0, 0};
        while (true) {
            // ================= ADVANCE =================
            if (a.dbg)
                c_adv_lanes += __popc(__ballot_sync(SONICS_FULL, mode == M_ADV));
            if (mode == M_ADV) {
                int req = 0; // 1: binomial(dn, dp)   2: poisson(lam)
                float dp = 0.0f;
                double lam = 0.0;
                while (req == 0 && mode == M_ADV) {
                    if (a.dbg && lane == (__ffs(__activemask()) - 1))
                        ++c_adv_trips;
                    if (pc == PC_RESULT) {
                        if (stage == 0) {
                            amp = x;
                            if (n_stage == 1) { // below the floor: plain amplification (:477-479)
                                cur += amp;
                                down = 0;
                                pc = PC_FINISH;
                            } else {
                                stage = 1;
                                dn = amp;
                                dp = p_slip;
                                req = 1;
                            }
                        } else if (stage == 1) {
                            slip = x;
                            stage = 2;
                            dn = slip;
                            dp = p_upn;
                            req = 1;
                        } else {
                            down = slip - x; // mol_down = mol_slip - mol_up
                            cur += amp - slip;
                            if (down > 0)
                                carry = x; // sic: up-stutter only lands when mol_down > 0 (:473-476)
                            pc = PC_FINISH;
                        }
                    } else if (pc == PC_FINISH) {
                        if (b > 0)
                            hist[(b - 1) * 32] = prev + down;
                        if (down > 0 && b - 1 < live_lo)
                            live_lo = b - 1;
                        if (cur != 0 && b > live_hi)
                            live_hi = b;
                        prev = cur;
                        down = 0;
                        ++b;
                        E_up = fmaf(E_up + 1.0f, e1_up, E_up); // expm1((al+1) L) from expm1(al L)
                        E_dn = fmaf(E_dn + 1.0f, e1_dn, E_dn);
                        pc = PC_BINSCAN;
                    } else if (pc == PC_BINSCAN) {
                        if (b > bend) {
                            hist[bend * 32] = prev;
                            ++cycle;
                            pc = PC_CYCLE;
                        } else {
                            const int raw = hist[b * 32];
                            cur = raw + carry;
                            carry = 0;
                            if (raw != 0) { // in the start-of-cycle snapshot (:424)
                                const bool stutter = b >= floor_bin;
                                if (stutter) {
                                    ct = cur;
                                    const float prob_up = -E_up, prob_down = -E_dn;
                                    p_slip = fmaf(-prob_up, prob_down, prob_up + prob_down); // 1-(1-pd)(1-pu)
                                    p_upn = __fdividef(prob_up, prob_up + prob_down);
                                }
                                n_stage = stutter ? 3 : 1;
                                stage = 0;
                                dn = ct;
                                dp = eff;
                                req = 1;
                                pc = PC_RESULT;
                            } else {
                                pc = PC_FINISH;
                            }
                        }
                    } else if (pc == PC_CYCLE) {
                        if (cycle > n_cycles) {
                            mode = M_DONE;
                        } else if (cycle == cap_cycle && !cap_done) {
                            // capture step (:412-423): max count, first non-zero allele, non-zero count
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
                            cap_done = true;
                            if (nnz > 0) {
                                cap_set = (float)((double)capture / (double)(maxp - minidx));
                                floor_cap = (float)(1.0 - ((double)__fmul_rn(cap_set, (float)nnz)) / 2.0);
                                cap_n = 1;
                                cap_b = live_lo;
                                cap_hi = live_hi;
                                pc = PC_CAPSCAN;
                            }
                        } else {
                            // begin the amplification sweep (:424-479)
                            b = live_lo;
                            bend = min(live_hi + 1, W - 1);
                            carry = 0;
                            down = 0;
                            prev = b > 0 ? hist[(b - 1) * 32] : 0;
                            const float al = (float)(lo + b);
                            E_up = expm1f(al * l_up);
                            E_dn = expm1f(al * l_dn);
                            pc = PC_BINSCAN;
                        }
                    } else if (pc == PC_CAPSCAN) {
                        if (cap_b > cap_hi) {
                            pc = PC_CYCLE;
                        } else {
                            cap_v = hist[cap_b * 32];
                            if (cap_v != 0) {
                                ct = cap_v;
                                const float hit = __fmul_rn(
                                    __fadd_rn(floor_cap, __fmul_rn(cap_set, (float)cap_n)), (float)cap_v);
                                lam = (double)hit;
                                pc = PC_CAPRESULT;
                                if (lam > 0.0)
                                    req = 2;
                                else
                                    x = 0;
                            } else {
                                ++cap_b;
                            }
                        }
                    } else { // PC_CAPRESULT
                        hist[cap_b * 32] = x < cap_v ? x : cap_v;
                        ++cap_n;
                        ++cap_b;
                        pc = PC_CAPSCAN;
                    }
                    if (req == 1) { // trivial draws never leave ADVANCE
                        if (dn <= 0 || !(dp > 0.0f)) {
                            x = 0;
                            req = 0;
                        } else if (dp >= 1.0f) {
                            x = dn;
                            req = 0;
                        }
                    }
                }
                // ---- set the draw up (every lane that leaves ADVANCE with a request is here) ----
                if (req == 1) {
                    flip = dp > 0.5f;
                    const float pp = flip ? 1.0f - dp : dp;
                    if ((float)dn * pp < 10.0f) {
                        inv_setup(iv, dn, pp);
                        mode = M_INV0;
                    } else {
                        btrs_setup(bs, dn, pp);
                        mode = M_BTRS;
                    }
                } else if (req == 2) {
                    pois_setup(ps, lam);
                    mode = M_POIS;
                }
            }
            if (__all_sync(SONICS_FULL, mode == M_DONE))
                break;
            if (a.dbg) {
                ++c_pass;
#pragma unroll
                for (int mm = 0; mm < 7; ++mm)
                    c_mode[mm] += __popc(__ballot_sync(SONICS_FULL, mode == mm));
            }

            // ================= one step of the draw in flight =================
            // A lane that starts an inversion or runs a rejection trial in this pass takes the next
            // 128-bit block of ITS OWN stream (the block counter advances only when the lane consumes,
            // never with the warp's pass count: a lane's words must not depend on how long it waited
            // for its neighbours).  Both pairs are used within the pass: the second one feeds an
            // immediate retry when the first trial is rejected outright.
            if (mode == M_INV0 || mode == M_BTRS || mode == M_POIS) {
                const uint4 words = rng.block4();
                if (mode == M_INV0) {
                    inv_start(iv, words.x);
                    mode = M_INV;
                } else if (mode == M_BTRS) {
                    int r = btrs_trial(bs, words.x, words.y);
                    if (r == 0)
                        r = btrs_trial(bs, words.z, words.w);
                    if (r == 1) {
                        const int k = (int)bs.fk;
                        x = flip ? dn - k : k;
                        mode = M_ADV;
                    } else if (r == 2) {
                        mode = M_FULL;
                    }
                } else {
                    if (ps.lam < 10.0) {
                        x = poisson_small(ps.lam, words.x);
                        mode = M_ADV;
                    } else if (pois_trial(ps, words.x, words.y) || pois_trial(ps, words.z, words.w)) {
                        x = ps.k;
                        mode = M_ADV;
                    }
                }
            }
            if (mode == M_INV) {
                if (inv_run(iv, rcp_tab, 8)) {
                    x = flip ? dn - iv.x : iv.x;
                    mode = M_ADV;
                }
            }

            // ================= deferred Stirling test =================
            {
                const unsigned m_full = __ballot_sync(SONICS_FULL, mode == M_FULL);
                const unsigned m_idle = __ballot_sync(SONICS_FULL, mode == M_FULL || mode == M_DONE);
                if (__popc(m_full) >= 4 || (m_full != 0u && m_idle == SONICS_FULL)) {
                    if (a.dbg) {
                        ++c_full_runs;
                        c_full_lanes += __popc(m_full);
                    }
                    if (mode == M_FULL) {
                        if (btrs_full(bs)) {
                            const int k = (int)bs.fk;
                            x = flip ? dn - k : k;
                            mode = M_ADV;
                        } else {
                            mode = M_BTRS;
                        }
                    }
                }
            }
        }

        if (a.dbg) {
            // c_adv_trips was counted by whichever lane led each trip: sum over the warp
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
                atomicAdd(&a.dbg[K1_CNT_ITEMS], 1ull);
            }
        }

        // ---- lnL (:360-363, pymc_extracted.f:88-108) ----
        const int wlo = min(__reduce_min_sync(SONICS_FULL, live_lo), in_lo);
        const int whi = max(__reduce_max_sync(SONICS_FULL, live_hi), in_hi);
        bool sentinel = false;
        int T = 0;
        for (int b = wlo; b <= whi; ++b) {
            const int c = hist[b * 32];
            sentinel |= (s_reads[b] > c);
            T += c;
        }
        double like = SONICS_SENTINEL;
        if (!sentinel) {
            like = 0.0;
            for (int b = wlo; b <= whi; ++b) {
                const int c = hist[b * 32];
                if (c == 0)
                    continue; // x == 0 too: the reference adds exact zeros
                const int x = s_reads[b];
                const double fc = factln_dev(c);
                like = __dadd_rn(like, fc);
                like = __dsub_rn(like, x ? s_fl[b] : 0.0);
                like = __dsub_rn(like, x ? factln_dev(c - x) : fc);
            }
            like = __dsub_rn(like, __dsub_rn(__dsub_rn(factln_dev(T), h.fl_D), factln_dev(T - D)));
        }
        const int64_t o = (int64_t)g * a.out_stride + a.out_offset + (j - a.sim_begin);
        if (valid && !a.replay) {
            a.lnL[o] = like;
            const int glo = min(i1, i2), ghi = max(i1, i2);
            a.group[o] = (uint16_t)(glo * A + ghi);
            if (a.simpar) {
                a.simpar[o * 4 + 0] = par_down;
                a.simpar[o * 4 + 1] = par_up;
                a.simpar[o * 4 + 2] = par_cap;
                a.simpar[o * 4 + 3] = par_eff;
            }
        }

        if (READOUT) {
            // ---- sequencing readout + descriptors (:341-357, :365-384, rsq :273-291) ----
            // own RNG stream: the main stream (and therefore lnL) is identical with or without it
            Rng rr;
            rr.init(a.seed, (uint64_t)j, h.uid, STREAM_READOUT);
            // stage 1: n_i ~ Binomial(D, pool_i / T) for every bin (:354)
            int N = 0;
            const double invT = T > 0 ? 1.0 / (double)T : 0.0;
            for (int b = wlo; b <= whi; ++b) {
                const int c = hist[b * 32];
                int ni = 0;
                if (c != 0) {
                    const uint4 w = rr.block4();
                    ni = binomial(rr, D, (float)((double)c * invT), w.x, w.y, rcp_tab);
                }
                hist[b * 32] = ni;
                N += ni;
            }
            double ident = 0.0;
            float r2 = -999999.0f;
            if (N > 0) {
                // stage 2: readout = bincount(choice(mid, D)) ~ Multinomial(D; n_i / N) (:368),
                // drawn as a chain of conditional binomials
                int R = D, M = N, fr = -1, to = -1, acc = 0;
                for (int b = wlo; b <= whi; ++b) {
                    const int ni = hist[b * 32];
                    int xi = 0;
                    if (ni > 0 && R > 0) {
                        const uint4 w = rr.block4();
                        xi = (ni == M) ? R : binomial(rr, R, __fdividef((float)ni, (float)M), w.x, w.y, rcp_tab);
                    }
                    R -= xi;
                    M -= ni;
                    hist[b * 32] = xi;
                    const int rd = s_reads[b];
                    acc += min(rd, xi);
                    if (rd != 0 || xi != 0) {
                        if (fr < 0)
                            fr = b;
                        to = b;
                    }
                }
                ident = (double)acc / (double)D;
                const double mean = (double)D / (double)(to - fr + 1);
                double ss_tot = 0.0;
                long long ss_res = 0;
                for (int b = fr; b <= to; ++b) {
                    const int rd = s_reads[b], xi = hist[b * 32];
                    const double dl = (double)rd - mean;
                    ss_tot = __dadd_rn(ss_tot, __dmul_rn(dl, dl));
                    const long long e = (long long)rd - xi;
                    ss_res += e * e;
                }
                if (ss_tot == 0.0)
                    ss_tot = 1e-16;
                r2 = (float)(1.0 - (double)ss_res / ss_tot);
            }
            if (valid && a.replay) {
                a.replay[g].identity = ident;
                a.replay[g].r_squared = (double)r2;
                if (a.replay[g].lnL != like)
                    a.replay[g].best_sim = -2; // replay diverged from the recorded simulation: surfaced by the host
            } else if (valid) {
                if (a.identity)
                    a.identity[o] = ident;
                if (a.rsq)
                    a.rsq[o] = r2;
            }
        }
    }
}

} // namespace sonics
