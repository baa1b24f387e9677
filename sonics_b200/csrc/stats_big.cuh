// K2/K3 for large simulation counts (n_run > SONICS_STATS_MAX_N): the same algorithm as
// stats_kernel.cuh, one genotype at a time, with the sort / scans done by CUB over global memory.
//
//   DeviceRadixSort::SortPairs   lnL (fp64 keys, ascending) with the simulation index as payload
//   k_big_slots                  group slot of every sorted element; per-slot count / valid count / max
//   DeviceSelect::Flagged        start index of every run of equal lnL (tie blocks of any union of groups)
//   DeviceScan::ExclusiveSum     S_g[i] = #{ j < i : slot_j == g }  (one per group)
//   k_big_pair                   per tie block: U += t_best (S_g[head] + t_g / 2),  tie += (t_best+t_g)^3 - (t_best+t_g)
//   k_big_pick                   order statistics of a group (quantiles)
//   k_big_verdict                top-2 groups, ratios, Bonferroni, PASS / FILTER  (same arithmetic as k_evaluate)
// Everything stays on the device; the host only enqueues.  Used for --monte_carlo with a large -r
// (sonics.pyx:109-134) and for -r above 8192 in the default mode.
#pragma once
#include <cub/cub.cuh>

#include "stats_kernel.cuh"

namespace sonics {

struct BigState { // device-resident scalars shared by the kernels of one genotype
    int G, min_sim, best, second, best_sim, evaluated;
    int n_heads;
    int pad;
    double high;        // max p so far
    int n_tests, pad2;
    double U;           // per-pair accumulators
    unsigned long long tie; // sum over tie blocks of t^3 - t, exact while (n1 + n2)^3 < 2^64
    double tie_d;           // the same sum in fp64 (scipy's tiecorrect works in float64) for larger samples
    double q[4];        // order statistics scratch
    double pct_ratio, best_ratio;
    double add_ratio[SONICS_MAX_ADD_RATIOS];
    double q75_lnL, q75_ident, q75_rsq;
};

struct BigBufs {
    double *keys;        // n   sorted lnL
    uint32_t *idx;       // n   simulation index of sorted element
    uint32_t *idx_in;    // n   iota
    uint16_t *slot;      // n   slot of sorted element
    unsigned char *flag; // n   run heads
    uint32_t *heads;     // n + 1
    uint32_t *S_best;    // n + 1
    uint32_t *S;         // n + 1
    double *tmp_in;      // n   (identity / r^2 of the best group, MC mode)
    double *tmp_out;     // n
    int *cnt;            // n_slots
    int *cnt_valid;      // n_slots
    unsigned long long *maxk; // n_slots
    BigState *st;
    void *cub_tmp;
    size_t cub_bytes;
};

__host__ inline size_t big_align(size_t v) { return (v + 255) & ~(size_t)255; }

__host__ inline size_t stats_big_bytes(int64_t n, int n_slots)
{
    size_t b = 0;
    b += big_align((size_t)n * 8);           // keys
    b += big_align((size_t)n * 4) * 2;       // idx, idx_in
    b += big_align((size_t)n * 2);           // slot
    b += big_align((size_t)n);               // flag
    b += big_align((size_t)(n + 1) * 4) * 3; // heads, S_best, S
    b += big_align((size_t)n * 8) * 2;       // tmp_in, tmp_out
    b += big_align((size_t)n_slots * 16);    // cnt, cnt_valid, maxk
    b += big_align(sizeof(BigState));
    b += big_align((size_t)n * 16 + (1 << 20)); // CUB temporary storage (radix sort of 8+4 B pairs dominates)
    return b;
}

struct SlotIs {
    const uint16_t *slot;
    int target;
    __host__ __device__ uint32_t operator()(uint32_t i) const { return slot[i] == target ? 1u : 0u; }
};
struct SlotIsDev { // target read from the device-resident state (best / second)
    const uint16_t *slot;
    const BigState *st;
    int which; // 0 best, 1 second, 2 best group of MC mode (= best)
    __device__ uint32_t operator()(uint32_t i) const { return slot[i] == (which == 1 ? st->second : st->best) ? 1u : 0u; }
};

__global__ void k_big_init(BigBufs b, int n, int n_slots)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n)
        b.idx_in[i] = (uint32_t)i;
    if (i < n_slots) {
        b.cnt[i] = 0;
        b.cnt_valid[i] = 0;
        b.maxk[i] = 0ull;
    }
    if (i == 0) {
        BigState s;
        memset(&s, 0, sizeof(s));
        s.best = s.second = s.best_sim = -1;
        *b.st = s;
    }
}

__global__ void k_big_slots(BigBufs b, const uint16_t *group, int n, int A, int first_idx, int random_start)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n)
        return;
    const int s = slot_of(group[b.idx[i]], A, first_idx, random_start);
    b.slot[i] = (uint16_t)s;
    const double x = b.keys[i];
    b.flag[i] = (i == 0 || x != b.keys[i - 1]) ? 1 : 0;
    atomicAdd(&b.cnt[s], 1);
    if (x > SONICS_SENTINEL)
        atomicAdd(&b.cnt_valid[s], 1);
    atomicMax(&b.maxk[s], f64_key(x));
}

// groups present, min_sim, top two by max lnL; closes the list of run heads with n
__global__ void k_big_top2(BigBufs b, int n, int n_slots, int min_sim_opt, const int *n_heads_dev)
{
    if (threadIdx.x != 0 || blockIdx.x != 0)
        return;
    BigState &st = *b.st;
    int G = 0, min_sim = 0x7fffffff, best = -1, second = -1;
    for (int s = 0; s < n_slots; ++s) {
        if (b.cnt[s] == 0)
            continue;
        ++G;
        min_sim = min(min_sim, b.cnt_valid[s]);
        if (best < 0 || b.maxk[s] > b.maxk[best]) {
            second = best;
            best = s;
        } else if (second < 0 || b.maxk[s] > b.maxk[second]) {
            second = s;
        }
    }
    st.G = G;
    st.min_sim = G ? min_sim : 0;
    st.best = best;
    st.second = second;
    st.evaluated = (G >= 2 && st.min_sim >= min_sim_opt) ? 1 : 0;
    st.n_heads = *n_heads_dev;
    b.heads[st.n_heads] = (uint32_t)n;
    if (best >= 0 && second >= 0)
        st.best_ratio = __dsub_rn(key_f64(b.maxk[best]), key_f64(b.maxk[second]));
}

// lowest simulation index among the rows of slot `which` tied at that slot's maximum
// (which < 0: the global maximum, MC mode -- also fixes st.best to that row's slot)
__global__ void k_big_best_row(BigBufs b, int n, int mc)
{
    __shared__ unsigned int best_idx;
    if (threadIdx.x == 0)
        best_idx = 0xffffffffu;
    __syncthreads();
    BigState &st = *b.st;
    const unsigned long long top = mc ? f64_key(b.keys[n - 1]) : b.maxk[st.best < 0 ? 0 : st.best];
    // equal keys are contiguous at the end of the group's range; scan backwards from the end
    for (int i = n - 1 - threadIdx.x; i >= 0; i -= blockDim.x) {
        const unsigned long long k = f64_key(b.keys[i]);
        if (k < top)
            break;
        if (k == top && (mc || b.slot[i] == st.best))
            atomicMin(&best_idx, b.idx[i]);
    }
    __syncthreads();
    if (threadIdx.x == 0)
        st.best_sim = (int)best_idx;
}

__global__ void k_big_set_best_slot(BigBufs b, const uint16_t *group, int64_t base, int A, int first_idx, int random_start)
{
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        BigState &st = *b.st;
        st.best = slot_of(group[base + st.best_sim], A, first_idx, random_start);
    }
}

// out q[0], q[1] = order statistics k0, k1 = k0 + 1 (clamped) of the group whose scan is S
__global__ void k_big_pick(BigBufs b, const uint32_t *S, int n, int which, double qv, int out0)
{
    const BigState &st = *b.st;
    const int target = which == 1 ? st.second : st.best;
    if (target < 0)
        return;
    const int cnt = b.cnt[target];
    const double hq = __dmul_rn((double)(cnt - 1), qv);
    const int k0 = (int)floor(hq), k1 = min(k0 + 1, cnt - 1);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        if (b.slot[i] == target) {
            if ((int)S[i] == k0)
                b.st->q[out0] = b.keys[i];
            if ((int)S[i] == k1)
                b.st->q[out0 + 1] = b.keys[i];
        }
    }
}

// combine the picked order statistics: ratio = quantile(best) - quantile(second)
__global__ void k_big_ratio(BigBufs b, double qv, int dest /* -1: pct_ratio, else add_ratio[dest] */)
{
    if (threadIdx.x != 0 || blockIdx.x != 0)
        return;
    BigState &st = *b.st;
    if (st.best < 0 || st.second < 0)
        return;
    double qq[2];
    for (int w = 0; w < 2; ++w) {
        const int cnt = b.cnt[w == 0 ? st.best : st.second];
        const double hq = __dmul_rn((double)(cnt - 1), qv);
        const int k0 = (int)floor(hq);
        qq[w] = lerp_np(st.q[2 * w], st.q[2 * w + 1], __dsub_rn(hq, (double)k0));
    }
    const double r = __dsub_rn(qq[0], qq[1]);
    if (dest < 0)
        st.pct_ratio = r;
    else
        st.add_ratio[dest] = r;
}

// t^3 - t in fp64 with fixed rounding points (compared for equality across kernels)
__device__ __forceinline__ double cube_minus(unsigned long long t)
{
    const double d = (double)t;
    return __dsub_rn(__dmul_rn(__dmul_rn(d, d), d), d);
}

__global__ void k_big_pair_begin(BigBufs b)
{
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        b.st->U = 0.0;
        b.st->tie = 0ull;
        b.st->tie_d = 0.0;
    }
}

// one thread per tie block of the global order
__global__ void k_big_pair(BigBufs b, int target)
{
    const BigState &st = *b.st;
    if (!st.evaluated || target == st.best || b.cnt[target] == 0)
        return;
    double u = 0.0, tie_d = 0.0;
    unsigned long long tie = 0ull;
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < st.n_heads; k += gridDim.x * blockDim.x) {
        const uint32_t hd = b.heads[k], en = b.heads[k + 1];
        const unsigned long long tb = b.S_best[en] - b.S_best[hd];
        const unsigned long long tg = b.S[en] - b.S[hd];
        u += (double)tb * ((double)b.S[hd] + 0.5 * (double)tg);
        const unsigned long long t = tb + tg;
        tie += t * t * t - t; // wraps for t > 2.6e6: k_big_pair_end then reads tie_d
        tie_d += cube_minus(t);
    }
    // warp reduce, then one atomic per warp (U is a sum of multiples of 1/2 below 2^53: exact in any order)
    for (int o = 16; o > 0; o >>= 1) {
        u += __shfl_down_sync(0xffffffffu, u, o);
        tie += __shfl_down_sync(0xffffffffu, tie, o);
        tie_d += __shfl_down_sync(0xffffffffu, tie_d, o);
    }
    if ((threadIdx.x & 31) == 0) {
        atomicAdd(&b.st->U, u);
        atomicAdd(&b.st->tie, tie);
        atomicAdd(&b.st->tie_d, tie_d);
    }
}

__global__ void k_big_pair_end(BigBufs b, int target)
{
    if (threadIdx.x != 0 || blockIdx.x != 0)
        return;
    BigState &st = *b.st;
    if (!st.evaluated || target == st.best || b.cnt[target] == 0)
        return;
    const int n1 = b.cnt[st.best], n2 = b.cnt[target];
    const double N = (double)(n1 + n2);
    ++st.n_tests;
    // N^3 - N fits 64 bits below N = 2^21; above that the fp64 sums are compared (one tie block of N values gives
    // exactly the same expression on both sides)
    const bool exact = n1 + n2 < (1 << 21);
    const unsigned long long Nu = (unsigned long long)(n1 + n2);
    const double tie_sum = exact ? (double)st.tie : st.tie_d;
    const bool all_equal = exact ? st.tie == Nu * Nu * Nu - Nu : st.tie_d == cube_minus(Nu);
    if (all_equal) {
        st.high = 1.0; // all values identical: ValueError in scipy 1.5.2 -> high_pval = 1 (sonics.pyx:175-176)
        return;
    }
    const double mu = __ddiv_rn((double)((long long)n1 * n2), 2.0);
    const double s_in = __dsub_rn(N + 1.0, __ddiv_rn(tie_sum, __dmul_rn(N, N - 1.0)));
    const double sd = sqrt(__dmul_rn(__ddiv_rn((double)((long long)n1 * n2), 12.0), s_in));
    const double z = __ddiv_rn(__dsub_rn(__dsub_rn(st.U, mu), 0.5), sd);
    const double p = 0.5 * erfc(__ddiv_rn(z, 1.4142135623730951));
    st.high = fmax(p, st.high);
}

// MC mode: gather identity / r^2 of the best group (everything else +inf so it sorts to the end)
__global__ void k_big_gather(BigBufs b, const uint16_t *group, const double *ident, const float *rsq, int64_t base, int n,
                             int A, int first_idx, int random_start, int which)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n)
        return;
    const int target = b.st->best;
    double v = __longlong_as_double(0x7ff0000000000000ll);
    if (slot_of(group[base + i], A, first_idx, random_start) == target)
        v = which == 0 ? (ident ? ident[base + i] : 0.0) : (rsq ? (double)rsq[base + i] : 0.0);
    b.tmp_in[i] = v;
}

__global__ void k_big_q75_sorted(BigBufs b, int which)
{
    if (threadIdx.x != 0 || blockIdx.x != 0)
        return;
    BigState &st = *b.st;
    const int cnt = b.cnt[st.best];
    const double hq = (double)(cnt - 1) * 0.75;
    const int k0 = (int)floor(hq), k1 = min(k0 + 1, cnt - 1);
    const double v = lerp_np(b.tmp_out[k0], b.tmp_out[k1], hq - (double)k0);
    if (which == 0)
        st.q75_ident = v;
    else
        st.q75_rsq = v;
}

__global__ void k_big_verdict(BigBufs b, const EvalArgs a, int g, int n)
{
    if (threadIdx.x != 0 || blockIdx.x != 0)
        return;
    const BigState &st = *b.st;
    const GtHeader h = a.t.hdr[g];
    const int A = h.n_alleles;
    const int64_t base = (int64_t)g * a.out_stride;
    SonicsVerdict v;
    memset(&v, 0, sizeof(v));
    v.allele_lo = v.allele_hi = -1;
    v.best_sim = -1;
    v.run_reps = n;
    v.n_groups = st.G;
    v.min_sim = st.min_sim;
    if (a.monte_carlo) {
        const int gid = a.group[base + st.best_sim];
        v.allele_lo = a.t.allele[h.a_off + gid / A];
        v.allele_hi = a.t.allele[h.a_off + gid % A];
        v.best_sim = st.best_sim;
        v.lnL = a.lnL[base + st.best_sim];
        v.identity = a.identity ? a.identity[base + st.best_sim] : 0.0;
        v.r_squared = a.rsq ? (double)a.rsq[base + st.best_sim] : 0.0;
        const int cnt = b.cnt[st.best];
        const double hq = (double)(cnt - 1) * 0.75;
        v.lnL_q75 = lerp_np(st.q[0], st.q[1], hq - floor(hq));
        v.identity_q75 = st.q75_ident;
        v.r_squared_q75 = st.q75_rsq;
        a.verdict[g] = v;
        return;
    }
    if (!st.evaluated) {
        if (a.final_round)
            v.filter = SONICS_FILTER_NO_SUCCESS;
        a.verdict[g] = v;
        return;
    }
    const double high = __dmul_rn(st.high, (double)st.n_tests);
    const int gid = a.group[base + st.best_sim];
    v.allele_lo = a.t.allele[h.a_off + gid / A];
    v.allele_hi = a.t.allele[h.a_off + gid % A];
    v.best_sim = st.best_sim;
    v.lnL = a.lnL[base + st.best_sim];
    v.best_lnL_ratio = st.best_ratio;
    v.percentile_lnL_ratio = st.pct_ratio;
    v.high_pval = high;
    for (int i = 0; i < a.n_add; ++i)
        v.add_ratio[i] = st.add_ratio[i];
    const bool pass = high < a.padjust && st.pct_ratio > a.thr && st.best_ratio > a.thr;
    if (pass) {
        v.filter = st.min_sim < 25 ? SONICS_FILTER_NO_SUCCESS : SONICS_FILTER_PASS; // :184-191 then :208, see k_evaluate
    } else if (a.final_round) {
        if (st.min_sim < 25) {
            v.filter = SONICS_FILTER_NO_SUCCESS;
        } else {
            if (!(high < 1.0)) {
                v.high_pval = 1.0;
                v.p_clamped = 1;
            }
            const double pr = a.n_add > 0 ? v.add_ratio[a.n_add - 1] : st.pct_ratio;
            int f = 0;
            if (v.high_pval > a.padjust)
                f |= SONICS_FILTER_MWU_TEST;
            if (st.best_ratio < a.thr)
                f |= SONICS_FILTER_BEST_RATIO;
            if (pr < a.thr)
                f |= SONICS_FILTER_PERCENTILE_RATIO;
            v.filter = f;
        }
    }
    if (a.n_add > 0 && (pass || a.final_round))
        v.percentile_lnL_ratio = v.add_ratio[a.n_add - 1];
    a.verdict[g] = v;
}

} // namespace sonics
