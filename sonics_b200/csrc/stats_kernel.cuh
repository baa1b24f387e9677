// K2/K3: per-genotype statistics on the cumulative lnL arrays -- one CTA per genotype.
//
// Restates the reductions of monte_carlo() (sonics.pyx:136-191, :208-252, MC mode :115-131):
//   groupby(genotype)                    -> group slots
//   min_sim = min_g #(lnL > -999999)     -> :140
//   groups ranked by max lnL, top two    -> :144-155
//   q75(best) - q75(second)              -> :158-160   (pandas linear quantile, sentinels included)
//   mannwhitneyu(best, g, "greater")     -> :166-178   (scipy asymptotic: midranks, tie-corrected
//                                                       sigma, continuity 0.5, p = sf(z))
//   Bonferroni, PASS predicate           -> :181-191
//   FILTER bits / no_success             -> :208-252
//
// Layout.  All n_run lnL values of the genotype are sorted ONCE (bitonic, shared memory) as
// order-preserving u64 keys with the simulation index as payload.  Everything else is read off
// that one global order: the tie blocks of the concatenation of any two groups are runs of equal
// keys, so for a pair (best, g) an exclusive block scan S_g of the indicator "element belongs to g"
// gives, for every element x of best,  #(g < x) = S_g[head(x)]  and  #(g == x) = S_g[end(x)] -
// S_g[head(x)],  i.e.  U = sum_x #(g<x) + #(g==x)/2  and the tie term  sum_blocks (t^3 - t)  with
// t = (S_best + S_g)[end] - (S_best + S_g)[head].  Order statistics of a group (quantiles) are the
// elements whose in-group prefix count hits the wanted rank.  fp64 throughout; the arithmetic
// that decides PASS is written with explicit round-to-nearest intrinsics in the same operation
// order as scipy/numpy so the verdicts agree with the reference's on identical lnL arrays.
#pragma once
#include <stdint.h>

#include <type_traits>

#include "../../include/sonics_b200.h"
#include "sim_kernel.cuh"

namespace sonics {

#define SONICS_STATS_THREADS 256
#define SONICS_STATS_MAX_N 8192   // simulations per genotype handled by the shared-memory path
#define SONICS_STATS_MAX_SLOTS 2048

struct EvalArgs {
    TableView t;
    const int32_t *active;
    int32_t n_run;
    int32_t n_pad; // power of two >= n_run
    int64_t out_stride;
    const double *lnL;
    const uint16_t *group;
    const double *identity; // MC mode only (nullable)
    const float *rsq;       // MC mode only (nullable)
    int32_t final_round;
    int32_t min_sim_opt, monte_carlo, random_start, n_add;
    double padjust, thr;
    double add_q[SONICS_MAX_ADD_RATIOS];
    SonicsVerdict *verdict;
    // global-memory form (k_evaluate_t<true>): per-genotype arrays at stride n_run + 1, filled by the host path
    // (abi_stats.inl: keys / indices sorted by CUB); slot_base = first active-list position of this launch
    int32_t slot_base;
    unsigned long long *g_key; // sorted lnL keys
    uint32_t *g_idx;           // simulation index of the sorted element; reused as S_best
    uint32_t *g_S;
    uint32_t *g_thead, *g_tend;
    uint16_t *g_slot;
    int *g_cnt, *g_cnt_valid;  // per genotype x n_slots_cap, only when the groups do not fit shared memory
    unsigned long long *g_maxk;
};

__device__ __forceinline__ unsigned long long f64_key(double v)
{
    const unsigned long long b = (unsigned long long)__double_as_longlong(v);
    return (b & 0x8000000000000000ull) ? ~b : (b | 0x8000000000000000ull);
}
__device__ __forceinline__ double key_f64(unsigned long long k)
{
    const unsigned long long b = (k & 0x8000000000000000ull) ? (k & 0x7fffffffffffffffull) : ~k;
    return __longlong_as_double((long long)b);
}

// numpy's linear quantile on an ascending sequence: the two order statistics a = v[lo], b = v[lo+1]
// and the weight t = h - lo are combined as numpy's _lerp does.
__device__ __forceinline__ double lerp_np(double a, double b, double t)
{
    const double d = __dsub_rn(b, a);
    double r = __dadd_rn(a, __dmul_rn(d, t));
    if (t >= 0.5)
        r = __dsub_rn(b, __dmul_rn(d, __dsub_rn(1.0, t)));
    if (d == 0.0 || t == 0.0)
        r = a;
    return r;
}

template <typename HT>
struct StatsArrays {
    unsigned long long *key; // n_pad
    uint32_t *idx;           // n_pad   (simulation index; later reused as S_best, n_pad + 1 needs <= n_pad + pad)
    uint32_t *S;             // n_pad + 1
    uint16_t *slot;          // n_pad
    HT *thead;               // n_pad   (u16 on chip: n <= 8192; u32 in global memory)
    HT *tend;                // n_pad
    int *cnt;                // n_slots
    int *cnt_valid;          // n_slots
    unsigned long long *maxk; // n_slots
    double *red;             // 32 doubles scratch
    uint32_t *wsum;          // 32 scratch
};

__host__ __device__ inline size_t stats_smem_bytes(int n_pad, int n_slots)
{
    size_t b = 0;
    b += (size_t)n_pad * 8;           // key
    b += (size_t)(n_pad + 2) * 4;     // idx / S_best
    b += (size_t)(n_pad + 2) * 4;     // S
    b += (size_t)n_pad * 2 * 3;       // slot, thead, tend
    b = (b + 7) & ~(size_t)7;
    b += (size_t)n_slots * 8;         // maxk
    b += (size_t)n_slots * 4 * 2;     // cnt, cnt_valid
    b = (b + 7) & ~(size_t)7;
    b += 32 * 8 + 32 * 4 + 64;
    return b;
}

__device__ __forceinline__ double block_sum(double v, double *red)
{
    for (int o = 16; o > 0; o >>= 1)
        v += __shfl_down_sync(0xffffffffu, v, o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0)
        red[threadIdx.x >> 5] = v;
    __syncthreads();
    double s = 0.0;
    if (threadIdx.x < 32) {
        s = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.0;
        for (int o = 16; o > 0; o >>= 1)
            s += __shfl_down_sync(0xffffffffu, s, o);
        if (threadIdx.x == 0)
            red[0] = s;
    }
    __syncthreads();
    s = red[0];
    __syncthreads();
    return s;
}

// S[i] = #{ j < i : slot[j] == target }, i in [0, n]
__device__ __forceinline__ void scan_slot(const uint16_t *slot, int target, int n, uint32_t *S, uint32_t *wsum)
{
    const int T = blockDim.x;
    const int per = (n + T - 1) / T;
    const int lo = min(n, (int)threadIdx.x * per), hi = min(n, lo + per);
    uint32_t c = 0;
    for (int i = lo; i < hi; ++i)
        c += (slot[i] == target);
    uint32_t inc = c;
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t y = __shfl_up_sync(0xffffffffu, inc, o);
        if ((threadIdx.x & 31) >= o)
            inc += y;
    }
    __syncthreads();
    if ((threadIdx.x & 31) == 31)
        wsum[threadIdx.x >> 5] = inc;
    __syncthreads();
    uint32_t base = 0;
    for (int w = 0; w < (int)(threadIdx.x >> 5); ++w)
        base += wsum[w];
    uint32_t run = base + inc - c;
    for (int i = lo; i < hi; ++i) {
        S[i] = run;
        run += (slot[i] == target);
    }
    if (hi == n && lo < n)
        S[n] = run;
    if (n == 0 && threadIdx.x == 0)
        S[0] = 0;
    __syncthreads();
}

// k-th smallest (0-based) element of group `target`, given its scan S
template <typename HT>
__device__ __forceinline__ void pick_order_stats(const StatsArrays<HT> &m, const uint32_t *S, int target, int n, int k0,
                                                 int k1, double *out /* smem[2] */)
{
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        if (m.slot[i] == target) {
            if ((int)S[i] == k0)
                out[0] = key_f64(m.key[i]);
            if ((int)S[i] == k1)
                out[1] = key_f64(m.key[i]);
        }
    }
    __syncthreads();
}

__device__ __forceinline__ void bitonic_sort(unsigned long long *key, uint32_t *idx, int n_pad)
{
    for (int k = 2; k <= n_pad; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = threadIdx.x; i < n_pad; i += blockDim.x) {
                const int ixj = i ^ j;
                if (ixj > i) {
                    const unsigned long long a = key[i], b = key[ixj];
                    const bool up = (i & k) == 0;
                    // ties are broken by the payload so the order is total and deterministic
                    const bool gt = a > b || (a == b && idx[i] > idx[ixj]);
                    if (gt == up) {
                        key[i] = b;
                        key[ixj] = a;
                        const uint32_t t = idx[i];
                        idx[i] = idx[ixj];
                        idx[ixj] = t;
                    }
                }
            }
            __syncthreads();
        }
    }
}

// Tie blocks of a sorted key array in global memory: thead[i] = first index of the run of equal keys that holds i,
// tend[i] = one past its last index.  A forward max-scan of the run heads and a backward min-scan of the run ends,
// each thread owning a contiguous chunk (the sentinel block at the bottom can hold millions of elements, so the
// per-run loop of the on-chip form would leave it to one thread).  wsum: 32 words of shared memory.
__device__ __forceinline__ void tie_blocks_global(const unsigned long long *key, int n, uint32_t *thead, uint32_t *tend,
                                                  uint32_t *wsum)
{
    const int T = blockDim.x, w = threadIdx.x >> 5, l = threadIdx.x & 31;
    const int per = (n + T - 1) / T;
    const int lo = min(n, (int)threadIdx.x * per), hi = min(n, lo + per);
    // forward: running maximum of the head positions
    uint32_t run = 0;
    for (int i = lo; i < hi; ++i)
        if (i > 0 && key[i] != key[i - 1])
            run = (uint32_t)i;
    uint32_t inc = run;
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t y = __shfl_up_sync(0xffffffffu, inc, o);
        if (l >= o)
            inc = max(inc, y);
    }
    __syncthreads();
    if (l == 31)
        wsum[w] = inc;
    __syncthreads();
    uint32_t carry = 0;
    for (int k = 0; k < w; ++k)
        carry = max(carry, wsum[k]);
    const uint32_t prev = __shfl_up_sync(0xffffffffu, inc, 1);
    if (l > 0)
        carry = max(carry, prev);
    for (int i = lo; i < hi; ++i) {
        if (i > 0 && key[i] != key[i - 1])
            carry = (uint32_t)i;
        thead[i] = carry;
    }
    __syncthreads();
    // backward: running minimum of the end positions
    uint32_t rmin = (uint32_t)n;
    for (int i = hi - 1; i >= lo; --i)
        if (i + 1 < n && key[i + 1] != key[i])
            rmin = (uint32_t)(i + 1);
    uint32_t dec = rmin;
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t y = __shfl_down_sync(0xffffffffu, dec, o);
        if (l + o < 32)
            dec = min(dec, y);
    }
    __syncthreads();
    if (l == 0)
        wsum[w] = dec;
    __syncthreads();
    uint32_t cmin = (uint32_t)n;
    for (int k = w + 1; k < (T >> 5); ++k)
        cmin = min(cmin, wsum[k]);
    const uint32_t nxt = __shfl_down_sync(0xffffffffu, dec, 1);
    if (l < 31)
        cmin = min(cmin, nxt);
    for (int i = hi - 1; i >= lo; --i) {
        if (i + 1 < n && key[i + 1] != key[i])
            cmin = (uint32_t)(i + 1);
        tend[i] = cmin;
    }
    __syncthreads();
}

// k-th smallest (0-based) of the keys key_of(i), i in [0, n) with member(i): an 8-pass MSD radix select, the
// histogram of one byte per pass in shared memory (hist: 256 words, sh: 2 ints).  No sort, no scratch: the q75 of the
// identity / r^2 column of the best group in --monte_carlo mode when the simulations live in global memory.
template <typename Member, typename KeyOf>
__device__ __forceinline__ unsigned long long radix_select(Member member, KeyOf key_of, int n, int k,
                                                           unsigned int *hist, int *sh)
{
    unsigned long long prefix = 0ull, mask = 0ull;
    for (int shift = 56; shift >= 0; shift -= 8) {
        for (int i = threadIdx.x; i < 256; i += blockDim.x)
            hist[i] = 0u;
        __syncthreads();
        for (int i = threadIdx.x; i < n; i += blockDim.x) {
            if (!member(i))
                continue;
            const unsigned long long key = key_of(i);
            if ((key & mask) == prefix)
                atomicAdd(&hist[(unsigned int)(key >> shift) & 255u], 1u);
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            int cum = 0, b = 0;
            for (; b < 255; ++b) {
                if (cum + (int)hist[b] > k)
                    break;
                cum += (int)hist[b];
            }
            sh[0] = b;
            sh[1] = k - cum;
        }
        __syncthreads();
        prefix |= (unsigned long long)sh[0] << shift;
        mask |= 0xffull << shift;
        k = sh[1];
        __syncthreads();
    }
    return prefix;
}

__device__ __forceinline__ int slot_of(int gid, int A, int first_idx, int random_start)
{
    if (random_start)
        return gid; // lo * A + hi
    const int lo = gid / A, hi = gid - lo * A;
    return lo == first_idx ? hi : lo; // index of the randomly chosen second allele
}

// BIG = false: everything in shared memory (n_run <= SONICS_STATS_MAX_N, groups <= SONICS_STATS_MAX_SLOTS), the sort
//               included (bitonic).
// BIG = true:  the per-simulation arrays live in global memory (EvalArgs::g_*), already sorted by CUB; the group
//              counters stay in shared memory when they fit (smem_slots != 0).  The --monte_carlo branch takes the
//              q75 of the identity / r^2 columns with a radix select instead of two more sorts.
template <bool BIG>
__global__ void __launch_bounds__(BIG ? 1024 : SONICS_STATS_THREADS) k_evaluate_t(const EvalArgs a, int n_slots_cap,
                                                                                  int smem_slots)
{
    extern __shared__ __align__(16) unsigned char sm_raw[];
    typedef typename std::conditional<BIG, uint32_t, uint16_t>::type HT;
    const int slot_ix = a.slot_base + blockIdx.x;
    const int g = a.active ? a.active[slot_ix] : slot_ix;
    const GtHeader h = a.t.hdr[g];
    const int A = h.n_alleles;
    const int n = a.n_run, n_pad = BIG ? a.n_run : a.n_pad;
    const int n_slots = a.random_start ? A * A : A;
    const int64_t base = (int64_t)g * a.out_stride;

    StatsArrays<HT> m;
    {
        unsigned char *p = sm_raw;
        if (BIG) {
            const size_t o = (size_t)blockIdx.x * ((size_t)n + 1);
            m.key = a.g_key + o;
            m.idx = a.g_idx + o;
            m.S = a.g_S + o;
            m.slot = a.g_slot + o;
            m.thead = (HT *)(a.g_thead + o);
            m.tend = (HT *)(a.g_tend + o);
        } else {
            m.key = (unsigned long long *)p;
            p += (size_t)n_pad * 8;
            m.idx = (uint32_t *)p;
            p += (size_t)(n_pad + 2) * 4;
            m.S = (uint32_t *)p;
            p += (size_t)(n_pad + 2) * 4;
            m.slot = (uint16_t *)p;
            p += (size_t)n_pad * 2;
            m.thead = (HT *)p;
            p += (size_t)n_pad * 2;
            m.tend = (HT *)p;
            p += (size_t)n_pad * 2;
            p = (unsigned char *)(((uintptr_t)p + 7) & ~(uintptr_t)7);
        }
        if (!BIG || smem_slots) {
            m.maxk = (unsigned long long *)p;
            p += (size_t)n_slots_cap * 8;
            m.cnt = (int *)p;
            p += (size_t)n_slots_cap * 4;
            m.cnt_valid = (int *)p;
            p += (size_t)n_slots_cap * 4;
            p = (unsigned char *)(((uintptr_t)p + 7) & ~(uintptr_t)7);
        } else {
            const size_t o = (size_t)blockIdx.x * n_slots_cap;
            m.maxk = a.g_maxk + o;
            m.cnt = a.g_cnt + o;
            m.cnt_valid = a.g_cnt_valid + o;
        }
        m.red = (double *)p;
        p += 32 * 8;
        m.wsum = (uint32_t *)p;
    }
    __shared__ int sh_i[8];
    __shared__ double sh_d[8];

    SonicsVerdict v;
    v.allele_lo = v.allele_hi = -1;
    v.filter = 0;
    v.run_reps = n;
    v.min_sim = 0;
    v.best_sim = -1;
    v.n_groups = 0;
    v.p_clamped = 0;
    v.lnL = v.identity = v.r_squared = 0.0;
    v.high_pval = 0.0;
    v.best_lnL_ratio = v.percentile_lnL_ratio = 0.0;
    v.identity_q75 = v.r_squared_q75 = v.lnL_q75 = 0.0;
    for (int i = 0; i < SONICS_MAX_ADD_RATIOS; ++i)
        v.add_ratio[i] = 0.0;

    // ---- 1. load, per-slot counts and maxima ----
    for (int i = threadIdx.x; i < n_slots; i += blockDim.x) {
        m.cnt[i] = 0;
        m.cnt_valid[i] = 0;
        m.maxk[i] = 0ull;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < n_pad; i += blockDim.x) {
        if (i < n) {
            const double x = a.lnL[base + i];
            const unsigned long long k = f64_key(x);
            if (!BIG) {
                m.key[i] = k;
                m.idx[i] = (uint32_t)i;
            }
            const int s = slot_of(a.group[base + i], A, h.first_idx, a.random_start);
            atomicAdd(&m.cnt[s], 1);
            if (x > SONICS_SENTINEL)
                atomicAdd(&m.cnt_valid[s], 1);
            atomicMax(&m.maxk[s], k);
        } else {
            m.key[i] = 0xffffffffffffffffull;
            m.idx[i] = 0xffffffffu;
        }
    }
    __syncthreads();

    // ---- 2. one global sort by lnL (BIG: done by CUB before the launch; stable, so ties keep ascending indices
    // exactly as the payload tie-break of the bitonic network orders them) ----
    if (!BIG)
        bitonic_sort(m.key, m.idx, n_pad);

    // ---- 3. slot of every sorted element, tie blocks ----
    for (int i = threadIdx.x; i < n; i += blockDim.x)
        m.slot[i] = (uint16_t)slot_of(a.group[base + m.idx[i]], A, h.first_idx, a.random_start);
    __syncthreads();
    if (BIG) {
        tie_blocks_global(m.key, n, (uint32_t *)m.thead, (uint32_t *)m.tend, m.wsum);
    } else {
        for (int i = threadIdx.x; i < n; i += blockDim.x) {
            // runs of equal keys are short except for the sentinel block at the bottom
            if (i == 0 || m.key[i] != m.key[i - 1]) {
                int e = i + 1;
                while (e < n && m.key[e] == m.key[i])
                    ++e;
                for (int j = i; j < e; ++j) {
                    m.thead[j] = (HT)i;
                    m.tend[j] = (HT)e; // n <= 8192 < 65536
                }
            }
        }
        __syncthreads();
    }

    // ---- 4. groups present, min_sim, top two by max lnL (thread 0; a few dozen slots) ----
    if (threadIdx.x == 0) {
        int G = 0, min_sim = 0x7fffffff, best = -1, second = -1;
        for (int s = 0; s < n_slots; ++s) {
            if (m.cnt[s] == 0)
                continue; // pandas groupby: groups without rows do not exist
            ++G;
            min_sim = min(min_sim, m.cnt_valid[s]);
            if (best < 0 || m.maxk[s] > m.maxk[best]) {
                second = best;
                best = s;
            } else if (second < 0 || m.maxk[s] > m.maxk[second]) {
                second = s;
            }
        }
        sh_i[0] = G;
        sh_i[1] = G ? min_sim : 0;
        sh_i[2] = best;
        sh_i[3] = second;
    }
    __syncthreads();
    const int G = sh_i[0], min_sim = sh_i[1], best = sh_i[2], second = sh_i[3];
    v.n_groups = G;
    v.min_sim = min_sim;

    if (a.monte_carlo) {
        // ---- MC mode (:115-131): best row = global max lnL; q75 of ident, r^2, lnL within its group ----
        const int gbest = m.slot[n - 1];
        // the reported row: lowest simulation index among the rows tied at the maximum
        uint32_t bi = 0xffffffffu;
        for (int i = threadIdx.x; i < n; i += blockDim.x)
            if (m.key[i] == m.key[n - 1])
                bi = min(bi, m.idx[i]);
        for (int o = 16; o > 0; o >>= 1)
            bi = min(bi, __shfl_down_sync(0xffffffffu, bi, o));
        if ((threadIdx.x & 31) == 0)
            m.wsum[threadIdx.x >> 5] = bi;
        __syncthreads();
        if (threadIdx.x == 0) {
            uint32_t r = 0xffffffffu;
            for (int w = 0; w < (int)(blockDim.x >> 5); ++w)
                r = min(r, m.wsum[w]);
            sh_i[4] = (int)r;
        }
        __syncthreads();
        const int best_sim = sh_i[4];
        const int gid = a.group[base + best_sim];
        const int gslot = slot_of(gid, A, h.first_idx, a.random_start);
        (void)gbest;
        const int cnt = m.cnt[gslot];
        const double hq = (double)(cnt - 1) * 0.75;
        const int k0 = (int)floor(hq), k1 = min(k0 + 1, cnt - 1);
        const double tq = hq - (double)k0;
        scan_slot(m.slot, gslot, n, m.S, m.wsum);
        pick_order_stats(m, m.S, gslot, n, k0, k1, sh_d);
        const double lnL_q75 = lerp_np(sh_d[0], sh_d[1], tq);
        __syncthreads();
        double q_extra[2] = {0.0, 0.0};
        for (int which = 0; which < 2 && BIG; ++which) {
            __shared__ unsigned int sh_hist[256];
            auto member = [&](int i) { return slot_of(a.group[base + i], A, h.first_idx, a.random_start) == gslot; };
            auto key_of = [&](int i) {
                return f64_key(which == 0 ? (a.identity ? a.identity[base + i] : 0.0)
                                          : (a.rsq ? (double)a.rsq[base + i] : 0.0));
            };
            const double lo_v = key_f64(radix_select(member, key_of, n, k0, sh_hist, sh_i + 5));
            const double hi_v = k1 == k0 ? lo_v : key_f64(radix_select(member, key_of, n, k1, sh_hist, sh_i + 5));
            q_extra[which] = lerp_np(lo_v, hi_v, tq);
        }
        for (int which = 0; which < 2 && !BIG; ++which) {
            // sort the group's identity (which = 0) / r^2 (which = 1) values; everything else -> +inf
            for (int i = threadIdx.x; i < n_pad; i += blockDim.x) {
                unsigned long long k = 0xffffffffffffffffull;
                if (i < n && slot_of(a.group[base + i], A, h.first_idx, a.random_start) == gslot) {
                    const double x = which == 0 ? (a.identity ? a.identity[base + i] : 0.0)
                                                : (a.rsq ? (double)a.rsq[base + i] : 0.0);
                    k = f64_key(x);
                }
                m.key[i] = k;
                m.idx[i] = (uint32_t)i;
            }
            __syncthreads();
            bitonic_sort(m.key, m.idx, n_pad);
            q_extra[which] = lerp_np(key_f64(m.key[k0]), key_f64(m.key[k1]), tq);
            __syncthreads();
        }
        if (threadIdx.x == 0) {
            v.allele_lo = a.t.allele[h.a_off + gid / A];
            v.allele_hi = a.t.allele[h.a_off + gid % A];
            v.best_sim = best_sim;
            v.lnL = a.lnL[base + best_sim];
            v.identity = a.identity ? a.identity[base + best_sim] : 0.0;
            v.r_squared = a.rsq ? (double)a.rsq[base + best_sim] : 0.0;
            v.lnL_q75 = lnL_q75;
            v.identity_q75 = q_extra[0];
            v.r_squared_q75 = q_extra[1];
            v.filter = 0;
            a.verdict[g] = v;
        }
        return;
    }

    const bool evaluated = G >= 2 && min_sim >= a.min_sim_opt;
    if (!evaluated) {
        if (threadIdx.x == 0) {
            // :208 hard-codes 25; min_sim in [25, --min_sim) crashes the reference (unbound name) and is
            // reported as no_success here
            if (a.final_round)
                v.filter = SONICS_FILTER_NO_SUCCESS;
            a.verdict[g] = v;
        }
        return;
    }

    // ---- best row of the best group (:227): lowest simulation index among rows tied at its maximum ----
    {
        uint32_t bi = 0xffffffffu;
        for (int i = threadIdx.x; i < n; i += blockDim.x)
            if (m.slot[i] == best && m.key[i] == m.maxk[best])
                bi = min(bi, m.idx[i]);
        for (int o = 16; o > 0; o >>= 1)
            bi = min(bi, __shfl_down_sync(0xffffffffu, bi, o));
        if ((threadIdx.x & 31) == 0)
            m.wsum[threadIdx.x >> 5] = bi;
        __syncthreads();
        if (threadIdx.x == 0) {
            uint32_t r = 0xffffffffu;
            for (int w = 0; w < (int)(blockDim.x >> 5); ++w)
                r = min(r, m.wsum[w]);
            sh_i[4] = (int)r;
        }
        __syncthreads();
    }
    const int best_sim = sh_i[4];
    __syncthreads();

    // ---- 5. quantile differences best vs second (:158-160, :231-242) ----
    uint32_t *S_best = m.idx; // the payload is no longer needed: reuse it for the scan of the best group
    scan_slot(m.slot, best, n, S_best, m.wsum);
    scan_slot(m.slot, second, n, m.S, m.wsum);
    const int n1 = m.cnt[best], n_sec = m.cnt[second];
    double pct_ratio = 0.0;
    for (int qi = 0; qi <= a.n_add; ++qi) {
        const double q = qi == 0 ? 0.75 : a.add_q[qi - 1];
        double qv[2];
        for (int which = 0; which < 2; ++which) {
            const int cnt = which == 0 ? n1 : n_sec;
            const double hq = __dmul_rn((double)(cnt - 1), q);
            const int k0 = (int)floor(hq), k1 = min(k0 + 1, cnt - 1);
            pick_order_stats(m, which == 0 ? S_best : m.S, which == 0 ? best : second, n, k0, k1, sh_d);
            qv[which] = lerp_np(sh_d[0], sh_d[1], __dsub_rn(hq, (double)k0));
            __syncthreads();
        }
        const double r = __dsub_rn(qv[0], qv[1]);
        if (qi == 0)
            pct_ratio = r;
        else
            v.add_ratio[qi - 1] = r;
    }
    const double best_ratio = __dsub_rn(key_f64(m.maxk[best]), key_f64(m.maxk[second]));

    // ---- 6. one-sided MWU of best against every other group (:166-178) ----
    double high = 0.0;
    int n_tests = 0;
    for (int s = 0; s < n_slots; ++s) {
        if (s == best || m.cnt[s] == 0)
            continue;
        ++n_tests;
        scan_slot(m.slot, s, n, m.S, m.wsum);
        const int n2 = m.cnt[s];
        double u_part = 0.0, tie_part = 0.0;
        for (int i = threadIdx.x; i < n; i += blockDim.x) {
            const int hd = m.thead[i], en = m.tend[i];
            if (m.slot[i] == best) {
                const double below = (double)m.S[hd];
                const double eq = (double)(m.S[en] - m.S[hd]);
                u_part += below + 0.5 * eq;
            }
            if (hd == i) {
                const double t = (double)((S_best[en] - S_best[hd]) + (m.S[en] - m.S[hd]));
                tie_part += t * t * t - t;
            }
        }
        const double U1 = block_sum(u_part, m.red);      // exact: multiples of 1/2 below 2^53
        const double tie_sum = block_sum(tie_part, m.red); // exact integers
        double p;
        const double N = (double)(n1 + n2);
        if (tie_sum == N * N * N - N) {
            p = 1.0; // every value identical: scipy 1.5.2 raises ValueError, the reference sets high_pval = 1
            high = 1.0;
        } else {
            const double mu = __ddiv_rn((double)((long long)n1 * n2), 2.0);
            const double s_in = __dsub_rn(N + 1.0, __ddiv_rn(tie_sum, __dmul_rn(N, N - 1.0)));
            const double sd = sqrt(__dmul_rn(__ddiv_rn((double)((long long)n1 * n2), 12.0), s_in));
            const double z = __ddiv_rn(__dsub_rn(__dsub_rn(U1, mu), 0.5), sd);
            p = 0.5 * erfc(__ddiv_rn(z, 1.4142135623730951));
            high = fmax(p, high);
        }
    }
    high = __dmul_rn(high, (double)n_tests); // Bonferroni (:181)

    // ---- 7. verdict ----
    if (threadIdx.x == 0) {
        const int gid = a.group[base + best_sim];
        v.allele_lo = a.t.allele[h.a_off + gid / A];
        v.allele_hi = a.t.allele[h.a_off + gid % A];
        v.best_sim = best_sim;
        v.lnL = a.lnL[base + best_sim];
        v.best_lnL_ratio = best_ratio;
        v.percentile_lnL_ratio = pct_ratio;
        v.high_pval = high;
        const bool pass = high < a.padjust && pct_ratio > a.thr && best_ratio > a.thr;
        if (pass) {
            // the loop breaks on success (:184-191) and the hard-coded `min_sim < 25` test still follows (:208):
            // with --min_sim below 25 a passing genotype whose min_sim lies in [--min_sim, 25) ends as no_success
            v.filter = min_sim < 25 ? SONICS_FILTER_NO_SUCCESS : SONICS_FILTER_PASS;
        } else if (a.final_round) {
            if (min_sim < 25) {
                v.filter = SONICS_FILTER_NO_SUCCESS;
            } else {
                if (!(high < 1.0)) {
                    v.high_pval = 1.0; // :225, printed as the int 1
                    v.p_clamped = 1;
                }
                // --add_ratios overwrites the percentile ratio before FILTER and the column are produced (:236-238)
                const double pr = a.n_add > 0 ? v.add_ratio[a.n_add - 1] : pct_ratio;
                int f = 0;
                if (v.high_pval > a.padjust)
                    f |= SONICS_FILTER_MWU_TEST;
                if (best_ratio < a.thr)
                    f |= SONICS_FILTER_BEST_RATIO;
                if (pr < a.thr)
                    f |= SONICS_FILTER_PERCENTILE_RATIO;
                v.filter = f;
            }
        }
        if (pass && !(high < 1.0)) { // cannot happen (padjust < 1) but keep the clamp semantics uniform
            v.high_pval = 1.0;
            v.p_clamped = 1;
        }
        if (a.n_add > 0 && (pass || a.final_round))
            v.percentile_lnL_ratio = v.add_ratio[a.n_add - 1];
        a.verdict[g] = v;
    }
}

// active' = genotypes of `active` whose verdict is not final yet (order preserved); *n_out = their count.
// One block; runs between rounds of the repeat-until-pass loop.
__global__ void k_compact_active(const int32_t *active_in, int n_in, const SonicsVerdict *verdict, int32_t *active_out,
                                 int32_t *n_out)
{
    __shared__ int s_base;
    __shared__ int s_warp[32];
    if (threadIdx.x == 0)
        s_base = 0;
    __syncthreads();
    for (int start = 0; start < n_in; start += blockDim.x) {
        const int i = start + threadIdx.x;
        int g = -1, keep = 0;
        if (i < n_in) {
            g = active_in ? active_in[i] : i;
            // finished: passed, or passed with min_sim < 25 (no_success before the last round, see k_evaluate)
            keep = !(verdict[g].filter & (SONICS_FILTER_PASS | SONICS_FILTER_NO_SUCCESS));
        }
        const unsigned bal = __ballot_sync(0xffffffffu, keep);
        const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
        if (lane == 0)
            s_warp[w] = __popc(bal);
        __syncthreads();
        int off = s_base;
        for (int k = 0; k < w; ++k)
            off += s_warp[k];
        if (keep)
            active_out[off + __popc(bal & ((1u << lane) - 1))] = g;
        __syncthreads();
        if (threadIdx.x == 0) {
            int tot = 0;
            for (int k = 0; k < (int)(blockDim.x >> 5); ++k)
                tot += s_warp[k];
            s_base += tot;
        }
        __syncthreads();
    }
    if (threadIdx.x == 0)
        *n_out = s_base;
}

} // namespace sonics
