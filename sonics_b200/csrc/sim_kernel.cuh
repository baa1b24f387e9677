// K1: the fused PCR-stutter simulation kernel (one lane = one simulation).
//
// Restates, per simulation, the reference's
//   generate_params   sonics.pyx:33-45
//   one_repeat        sonics.pyx:293-394  (start alleles, noise, lnL, readout, identity, r^2)
//   simulate          sonics.pyx:397-480  (capture step, amplification / stutter cycles)
// and the lnL of pymc_extracted.f:63-109 (lnl.cuh).
//
// Mapping onto the GPU.  All 32 lanes of a warp simulate the SAME genotype
// (same reads, same live window), so the genotype description is staged once per
// warp in shared memory and every loop bound is warp-uniform.  The allele-length
// histogram of each simulation lives in shared memory as a column
// hist[bin * 32 + lane] (bank = lane: conflict free); one cycle is one ascending
// sweep over the live window with the running bin kept in registers:
//   raw   = hist[b]             start-of-cycle count  -> "nzp" snapshot  (:424)
//   cur   = raw + carry         count at visit time, includes the same-cycle
//                               up-stutter carried in from bin b-1            (:429)
//   amplify / slip / up  -> three chained binomials                           (:454-462)
//   down-stutter is added to bin b-1, which is still in a register            (:472)
//   up-stutter becomes `carry` for bin b+1, gated on mol_down > 0 (sic, :473) (:476)
// so a cycle costs one LDS + one STS per live bin on top of the samplers.
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
};

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
        } else { // idle lane of a partially filled warp: empty pool, skips every bin
            live_lo = W;
            live_hi = -1;
        }

        // ---- simulate (:397-480) ----
        int ct = 0; // function-scope `ct` of the reference; goes stale below the floor (:478)
        const int floor_bin = h.floor_al - lo;
        for (int cycle = 1; cycle <= a.p.n_cycles; ++cycle) {
            const int wlo = __reduce_min_sync(SONICS_FULL, live_lo);
            const int whi = __reduce_max_sync(SONICS_FULL, live_hi);
            if (whi < wlo)
                break; // warp with no valid lane
            if (cycle == a.p.capture_cycle) {
                // capture step (:412-423)
                int maxp = 0, nnz = 0, minidx = -1;
                for (int b = wlo; b <= whi; ++b) {
                    const int v = hist[b * 32];
                    maxp = max(maxp, v);
                    if (v != 0) {
                        ++nnz;
                        if (minidx < 0)
                            minidx = lo + b;
                    }
                }
                if (nnz > 0) {
                    const float cap_set = (float)((double)capture / (double)(maxp - minidx));
                    const float floor_cap =
                        (float)(1.0 - ((double)__fmul_rn(cap_set, (float)nnz)) / 2.0);
                    int n = 1;
                    for (int b = wlo; b <= whi; ++b) {
                        const int v = hist[b * 32];
                        if (v != 0) {
                            ct = v;
                            const float hit =
                                __fmul_rn(__fadd_rn(floor_cap, __fmul_rn(cap_set, (float)n)), (float)v);
                            int k = poisson(rng, (double)hit);
                            k = k < v ? k : v;
                            hist[b * 32] = k;
                            ++n;
                        }
                    }
                }
            }
            // amplification sweep (:424-479)
            const int bend = min(whi + 1, W - 1);
            int carry = 0;
            int prev = (wlo > 0) ? hist[(wlo - 1) * 32] : 0; // finished value of bin b-1
            for (int b = wlo; b <= bend; ++b) {
                const int raw = hist[b * 32];
                int cur = raw + carry;
                carry = 0;
                int down = 0;
                if (raw != 0) {
                    const bool stutter = b >= floor_bin;
                    if (stutter)
                        ct = cur;
                    // random words for the first trial of each chained draw, fetched at a point
                    // where every lane that visits this bin is active
                    const uint4 ra = rng.block4();
                    uint4 rb = make_uint4(0u, 0u, 0u, 0u);
                    float p_slip = 0.0f, p_upn = 0.0f;
                    if (stutter) {
                        rb = rng.block4();
                        const float al = (float)(lo + b);
                        const float prob_up = -expm1f(al * l_up);
                        const float prob_down = -expm1f(al * l_dn);
                        p_slip = fmaf(-prob_up, prob_down, prob_up + prob_down); // 1 - (1-pd)(1-pu)
                        p_upn = __fdividef(prob_up, prob_up + prob_down);
                    }
                    int n = ct;
                    float p = eff;
                    int amp = 0, slip = 0, upn = 0;
                    const int n_stage = stutter ? 3 : 1;
#pragma unroll 1
                    for (int stage = 0; stage < n_stage; ++stage) {
                        const uint32_t ua = stage == 0 ? ra.x : (stage == 1 ? ra.z : rb.x);
                        const uint32_t ub = stage == 0 ? ra.y : (stage == 1 ? ra.w : rb.y);
                        const int x = binomial(rng, n, p, ua, ub, rcp_tab);
                        if (stage == 0) {
                            amp = x;
                            p = p_slip;
                        } else if (stage == 1) {
                            slip = x;
                            p = p_upn;
                        } else {
                            upn = x;
                        }
                        n = x;
                    }
                    down = slip - upn;
                    cur += amp - slip;
                    if (down > 0)
                        carry = upn; // sic: up-stutter only lands when mol_down > 0 (:473-476)
                }
                if (b > 0)
                    hist[(b - 1) * 32] = prev + down;
                if (down > 0 && b - 1 < live_lo)
                    live_lo = b - 1;
                if (cur != 0 && b > live_hi)
                    live_hi = b;
                prev = cur;
            }
            hist[bend * 32] = prev;
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
