// extern "C" entry points for the statistics / repeat-until-pass path (included by sonics_b200.cu).

static size_t big_align(size_t v) { return (v + 255) & ~(size_t)255; }

static int next_pow2(int v)
{
    int p = 1;
    while (p < v)
        p <<= 1;
    return p;
}

static void fill_eval_args(EvalArgs &a, const SonicsParams *params, const TableView &tv, const int32_t *dev_active,
                           int64_t n_run, int64_t out_stride, const double *dev_lnL, const uint16_t *dev_group,
                           const double *dev_identity, const float *dev_rsq, int final_round,
                           SonicsVerdict *dev_verdict)
{
    a.t = tv;
    a.active = dev_active;
    a.n_run = (int32_t)n_run;
    a.n_pad = std::max(next_pow2((int)std::min<int64_t>(n_run, 1 << 30)), 2);
    a.out_stride = out_stride;
    a.lnL = dev_lnL;
    a.group = dev_group;
    a.identity = dev_identity;
    a.rsq = dev_rsq;
    a.final_round = final_round;
    a.min_sim_opt = params->min_sim;
    a.monte_carlo = params->monte_carlo;
    a.random_start = params->random_start;
    a.n_add = params->n_add_ratios;
    a.padjust = params->padjust;
    a.thr = params->lnL_threshold;
    for (int i = 0; i < SONICS_MAX_ADD_RATIOS; ++i)
        a.add_q[i] = i < params->n_add_ratios ? params->add_ratios[i] : 0.0;
    a.verdict = dev_verdict;
    a.slot_base = 0;
    a.g_key = nullptr;
    a.g_idx = a.g_S = a.g_thead = a.g_tend = nullptr;
    a.g_slot = nullptr;
    a.g_cnt = a.g_cnt_valid = nullptr;
    a.g_maxk = nullptr;
}

// ---- global-memory form of the default-mode statistics, many genotypes per launch -------------------------
// Scratch of one chunk of `c` genotypes, n simulations each (arrays at stride n + 1 per genotype):
//   keys x2 (u64, CUB double buffer; the spare one becomes thead / tend), indices x2 (u32; the spare one becomes S,
//   the sorted one is reused as S_best), slot (u16), per-genotype group counters when they do not fit shared
//   memory, CUB temporary storage.  26 bytes per simulation.
struct GlobalLayout {
    size_t off_k0, off_k1, off_v0, off_v1, off_slot, off_cnt, off_tmp, tmp_bytes, total;
};
struct SegBegin {
    int stride;
    __host__ __device__ int operator()(int i) const { return i * stride; }
};
struct SegEnd {
    int stride, n;
    __host__ __device__ int operator()(int i) const { return i * stride + n; }
};
typedef cub::TransformInputIterator<int, SegBegin, cub::CountingInputIterator<int>> SegBeginIt;
typedef cub::TransformInputIterator<int, SegEnd, cub::CountingInputIterator<int>> SegEndIt;

static bool slots_fit_smem(int n_slots) { return n_slots <= SONICS_STATS_MAX_SLOTS; }

static int global_layout(int64_t c, int64_t n, int n_slots, GlobalLayout &L)
{
    const size_t N = (size_t)c * (size_t)(n + 1);
    size_t o = 0;
    L.off_k0 = o;
    o += big_align(N * 8);
    L.off_k1 = o;
    o += big_align(N * 8);
    L.off_v0 = o;
    o += big_align(N * 4);
    L.off_v1 = o;
    o += big_align(N * 4);
    L.off_slot = o;
    o += big_align(N * 2);
    L.off_cnt = o;
    if (!slots_fit_smem(n_slots))
        o += big_align((size_t)c * n_slots * 16);
    L.off_tmp = o;
    size_t tmp = 0;
    cub::DoubleBuffer<unsigned long long> dk(nullptr, nullptr);
    cub::DoubleBuffer<uint32_t> dv(nullptr, nullptr);
    // the last chunk of a launch sequence may hold a single genotype, which is sorted by the plain radix sort:
    // the temporary storage covers both forms
    cudaError_t e = cub::DeviceRadixSort::SortPairs(nullptr, tmp, dk, dv, (int)n, 0, 64);
    if (e != cudaSuccess)
        return -1;
    if (c > 1) {
        size_t tmp_seg = 0;
        cub::CountingInputIterator<int> iota(0);
        SegBeginIt b(iota, SegBegin{(int)(n + 1)});
        SegEndIt en(iota, SegEnd{(int)(n + 1), (int)n});
        e = cub::DeviceSegmentedRadixSort::SortPairs(nullptr, tmp_seg, dk, dv, (int)N, (int)c, b, en, 0, 64);
        if (e != cudaSuccess)
            return -1;
        tmp = std::max(tmp, tmp_seg);
    }
    L.tmp_bytes = big_align(tmp) + 256;
    L.total = o + L.tmp_bytes;
    return 0;
}

__global__ void k_global_keys(const EvalArgs a, int n_chunk, unsigned long long *key, uint32_t *idx)
{
    const int64_t stride = (int64_t)a.n_run + 1;
    const int64_t total = (int64_t)n_chunk * stride;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(t / stride), j = (int)(t - (int64_t)c * stride);
        if (j >= a.n_run)
            continue;
        const int slot_ix = a.slot_base + c;
        const int g = a.active ? a.active[slot_ix] : slot_ix;
        key[t] = f64_key(a.lnL[(int64_t)g * a.out_stride + j]);
        idx[t] = (uint32_t)j;
    }
}

static int launch_evaluate_global(sonics_ctx *ctx, EvalArgs a, int n_active, int n_slots, void *scratch,
                                  size_t scratch_bytes, cudaStream_t s)
{
    const int64_t n = a.n_run;
    GlobalLayout L;
    if (global_layout(1, n, n_slots, L))
        return fail(SONICS_ERR_CUDA, "CUB temporary-storage query failed");
    if (!scratch || scratch_bytes < L.total)
        return fail(SONICS_ERR_SPACE, "statistics over %lld simulations per genotype need a scratch buffer of at least "
                                      "%zu bytes (sonics_stats_scratch_bytes), %zu given",
                    (long long)n, L.total, scratch_bytes);
    // genotypes per chunk: what the scratch holds, below 2^31 array elements
    int64_t chunk = std::min<int64_t>(n_active, ((int64_t)1 << 31) / (n + 1) - 1);
    chunk = std::max<int64_t>(chunk, 1);
    while (chunk > 1) {
        if (global_layout(chunk, n, n_slots, L))
            return fail(SONICS_ERR_CUDA, "CUB temporary-storage query failed");
        if (L.total <= scratch_bytes)
            break;
        chunk = std::max<int64_t>(1, (int64_t)((double)chunk * (double)scratch_bytes / (double)L.total * 0.97));
    }
    if (global_layout(chunk, n, n_slots, L) || L.total > scratch_bytes)
        return fail(SONICS_ERR_SPACE, "statistics scratch of %zu bytes does not hold one genotype", scratch_bytes);
    unsigned char *base = (unsigned char *)scratch;
    const bool smem_slots = slots_fit_smem(n_slots);
    const size_t smem = (smem_slots ? (size_t)n_slots * 16 + 8 : 0) + 32 * 8 + 32 * 4 + 64;
    CUDA_OK(cudaFuncSetAttribute(k_evaluate_t<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    for (int64_t c0 = 0; c0 < n_active; c0 += chunk) {
        const int c = (int)std::min<int64_t>(chunk, n_active - c0);
        const size_t N = (size_t)c * (size_t)(n + 1);
        a.slot_base = (int)c0;
        unsigned long long *k0 = (unsigned long long *)(base + L.off_k0), *k1 = (unsigned long long *)(base + L.off_k1);
        uint32_t *v0 = (uint32_t *)(base + L.off_v0), *v1 = (uint32_t *)(base + L.off_v1);
        const int T = 256, nb = (int)std::min<int64_t>((int64_t)((N + T - 1) / T), (int64_t)ctx->sm_count * 32);
        k_global_keys<<<nb, T, 0, s>>>(a, c, k0, v0);
        ctx->launches++;
        cub::DoubleBuffer<unsigned long long> dk(k0, k1);
        cub::DoubleBuffer<uint32_t> dv(v0, v1);
        size_t tmp = L.tmp_bytes;
        if (c == 1) {
            CUDA_OK(cub::DeviceRadixSort::SortPairs(base + L.off_tmp, tmp, dk, dv, (int)n, 0, 64, s));
        } else {
            cub::CountingInputIterator<int> iota(0);
            SegBeginIt b(iota, SegBegin{(int)(n + 1)});
            SegEndIt en(iota, SegEnd{(int)(n + 1), (int)n});
            CUDA_OK(cub::DeviceSegmentedRadixSort::SortPairs(base + L.off_tmp, tmp, dk, dv, (int)N, c, b, en, 0, 64, s));
        }
        ctx->launches++;
        a.g_key = dk.Current();
        a.g_idx = dv.Current();
        a.g_S = dv.Alternate();
        a.g_thead = (uint32_t *)dk.Alternate();
        a.g_tend = a.g_thead + N;
        a.g_slot = (uint16_t *)(base + L.off_slot);
        if (!smem_slots) {
            unsigned char *p = base + L.off_cnt;
            a.g_maxk = (unsigned long long *)p;
            a.g_cnt = (int *)(p + (size_t)c * n_slots * 8);
            a.g_cnt_valid = a.g_cnt + (size_t)c * n_slots;
        }
        k_evaluate_t<true><<<c, 1024, smem, s>>>(a, n_slots, smem_slots ? 1 : 0);
        ctx->launches++;
        CUDA_OK(cudaGetLastError());
    }
    return SONICS_OK;
}

static int launch_evaluate(sonics_ctx *ctx, const SonicsParams *params, const TableInfo &ti, const TableView &tv,
                           const int32_t *dev_active, int n_active, int64_t n_run, int64_t out_stride,
                           const double *dev_lnL, const uint16_t *dev_group, const double *dev_identity,
                           const float *dev_rsq, int final_round, SonicsVerdict *dev_verdict, void *dev_scratch,
                           size_t dev_scratch_bytes, void *stream)
{
    if (n_run <= 0 || n_run > 2147483000LL)
        return fail(SONICS_ERR_RANGE, "statistics over %lld simulations per genotype", (long long)n_run);
    if (params->n_add_ratios < 0 || params->n_add_ratios > SONICS_MAX_ADD_RATIOS)
        return fail(SONICS_ERR_ARG, "at most %d --add_ratios percentiles are supported", SONICS_MAX_ADD_RATIOS);
    const int n_slots = params->random_start ? ti.max_A * ti.max_A : ti.max_A;
    EvalArgs a;
    fill_eval_args(a, params, tv, dev_active, n_run, out_stride, dev_lnL, dev_group, dev_identity, dev_rsq, final_round,
                   dev_verdict);
    cudaStream_t s = (cudaStream_t)stream;
    if (n_run <= SONICS_STATS_MAX_N && n_slots <= SONICS_STATS_MAX_SLOTS) {
        const size_t smem = stats_smem_bytes(a.n_pad, n_slots);
        CUDA_OK(cudaFuncSetAttribute(k_evaluate_t<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_evaluate_t<false><<<n_active, SONICS_STATS_THREADS, smem, s>>>(a, n_slots, 1);
        ctx->launches++;
        CUDA_OK(cudaGetLastError());
        return SONICS_OK;
    }
    // more simulations per genotype than shared memory holds (-r above 8192), or more start-allele groups
    // (--random with more than 45 alleles): the same kernel over global-memory arrays sorted by CUB, many
    // genotypes per launch, both modes
    return launch_evaluate_global(ctx, a, n_active, n_slots, dev_scratch, dev_scratch_bytes, s);
}

extern "C" {

size_t sonics_stats_scratch_bytes_batch(int n_active, int64_t n_run, int max_alleles, int random_start)
{
    const int n_slots = random_start ? max_alleles * max_alleles : max_alleles;
    if ((n_run <= SONICS_STATS_MAX_N && n_slots <= SONICS_STATS_MAX_SLOTS) || n_run <= 0 || n_active <= 0)
        return 0;
    // all genotypes in one chunk (anything from the size for one genotype up works: more chunks)
    const int64_t c = std::max<int64_t>(1, std::min<int64_t>(n_active, ((int64_t)1 << 31) / (n_run + 1) - 1));
    GlobalLayout L;
    return global_layout(c, n_run, n_slots, L) == 0 ? L.total : 0;
}

size_t sonics_stats_scratch_bytes(int64_t n_run, int max_alleles, int random_start)
{
    return sonics_stats_scratch_bytes_batch(1, n_run, max_alleles, random_start);
}

int sonics_evaluate(sonics_ctx *ctx, const SonicsParams *params, const void *dev_table, int n_gt,
                    const int32_t *dev_active, int n_active, int64_t n_run, int64_t out_stride, const double *dev_lnL,
                    const uint16_t *dev_group, const double *dev_identity, const float *dev_rsq, int final_round,
                    SonicsVerdict *dev_verdict, void *dev_scratch, size_t dev_scratch_bytes, void *stream)
{
    if (!ctx || !params || !dev_table || !dev_lnL || !dev_group || !dev_verdict)
        return fail(SONICS_ERR_ARG, "sonics_evaluate: NULL argument");
    if (out_stride < n_run)
        return fail(SONICS_ERR_SPACE, "sonics_evaluate: out_stride < n_run");
    TableInfo ti;
    TableView tv;
    if (int rc = find_table(ctx, dev_table, n_gt, ti, tv))
        return rc;
    if (!dev_active)
        n_active = n_gt;
    if (n_active <= 0)
        return SONICS_OK;
    CUDA_OK(cudaSetDevice(ctx->device));
    return launch_evaluate(ctx, params, ti, tv, dev_active, n_active, n_run, out_stride, dev_lnL, dev_group, dev_identity,
                           dev_rsq, final_round, dev_verdict, dev_scratch, dev_scratch_bytes, stream);
}

int sonics_replay_best(sonics_ctx *ctx, const SonicsParams *params, const void *dev_table, int n_gt,
                       const int32_t *dev_active, int n_active, uint64_t seed, SonicsVerdict *dev_verdict,
                       void *dev_scratch, size_t dev_scratch_bytes, void *stream)
{
    if (!ctx || !dev_table || !dev_verdict)
        return fail(SONICS_ERR_ARG, "sonics_replay_best: NULL argument");
    if (int rc = check_params(params))
        return rc;
    TableInfo ti;
    TableView tv;
    if (int rc = find_table(ctx, dev_table, n_gt, ti, tv))
        return rc;
    if (!dev_active)
        n_active = n_gt;
    if (n_active <= 0)
        return SONICS_OK;
    CUDA_OK(cudaSetDevice(ctx->device));
    SimArgs a;
    a.t = tv;
    a.p = dev_params(params);
    a.active = dev_active;
    a.sims_per_gt = 1;
    a.sim_begin = 0;
    a.seed = seed;
    a.out_stride = 0;
    a.out_offset = 0;
    a.lnL = nullptr;
    a.group = nullptr;
    a.identity = nullptr;
    a.rsq = nullptr;
    a.simpar = nullptr;
    a.replay = dev_verdict;
    return launch_simulate(ctx, a, ti, true, n_active, n_active, dev_scratch, dev_scratch_bytes, stream);
}

// device workspace of sonics_genotype_batch
struct BatchLayout {
    size_t off_table, off_lnL, off_group, off_identity, off_rsq, off_verdict, off_act0, off_act1, off_count,
        off_scratch, scratch_bytes, total;
    int64_t stride;
};

// n_slots_max: largest number of start-allele groups of a genotype (A, or A * A with --random); 0 = not known,
// assumed to fit the on-chip statistics
static BatchLayout batch_layout(int n_gt, int n_alleles_total, int max_reps, int monte_carlo, int n_slots_max = 0)
{
    BatchLayout L;
    size_t o = 0;
    L.stride = max_reps;
    L.off_table = o;
    o += align_up(sonics_table_bytes(n_gt, n_alleles_total), 256);
    L.off_lnL = o;
    o += align_up((size_t)n_gt * L.stride * 8, 256);
    L.off_group = o;
    o += align_up((size_t)n_gt * L.stride * 2, 256);
    L.off_identity = o;
    if (monte_carlo)
        o += align_up((size_t)n_gt * L.stride * 8, 256);
    L.off_rsq = o;
    if (monte_carlo)
        o += align_up((size_t)n_gt * L.stride * 4, 256);
    L.off_verdict = o;
    o += align_up((size_t)n_gt * sizeof(SonicsVerdict), 256);
    L.off_act0 = o;
    o += align_up((size_t)n_gt * 4, 256);
    L.off_act1 = o;
    o += align_up((size_t)n_gt * 4, 256);
    L.off_count = o;
    o += 256;
    // simulation scratch: sized so that the largest round of the schedule (sonics.pyx:79-82,193-199) runs as one
    // chunk of the lockstep form (24 bytes per work unit + CUB storage), never below the smallest chunk of the
    // persistent form for the widest possible window; smaller scratch would only mean more chunks.
    L.off_scratch = o;
    int64_t run = 0, biggest = 0, reps = monte_carlo ? max_reps : (max_reps > 100 ? 100 : max_reps);
    for (int rnd = 1; run < max_reps; ++rnd) {
        biggest = std::max(biggest, reps);
        run += reps;
        reps = (rnd % 2 == 1) ? 4 * run : run;
        if (reps + run > max_reps)
            reps = max_reps - run;
    }
    LsLayout ls;
    const int64_t units = std::min<int64_t>((int64_t)n_gt * biggest, (int64_t)1 << 30);
    L.scratch_bytes = scratch_bytes_for(32 * 8, SONICS_MAX_ALLELE);
    if (ls_layout(units, ls) == 0)
        L.scratch_bytes = std::max(L.scratch_bytes, ls.total);
    // the large-n statistics reuse the scratch between rounds: up to 8 GiB of it (fewer genotypes per statistics
    // launch beyond that), never less than one genotype needs
    if (max_reps > SONICS_STATS_MAX_N || n_slots_max > SONICS_STATS_MAX_SLOTS) {
        GlobalLayout g1, gall;
        const int ns = std::max(n_slots_max, 1);
        const int64_t c = std::max<int64_t>(1, std::min<int64_t>(n_gt, ((int64_t)1 << 31) / ((int64_t)max_reps + 1) - 1));
        size_t one = 0, all = 0;
        if (global_layout(1, max_reps, ns, g1) == 0)
            one = g1.total;
        if (global_layout(c, max_reps, ns, gall) == 0)
            all = gall.total;
        L.scratch_bytes = std::max(L.scratch_bytes, std::max(one, std::min<size_t>(all, (size_t)8 << 30)));
    }
    o += align_up(L.scratch_bytes, 256);
    L.total = o;
    return L;
}

size_t sonics_genotype_work_bytes_mode(int n_gt, int n_alleles_total, int max_reps, int monte_carlo)
{
    if (n_gt <= 0 || n_alleles_total <= 0 || max_reps <= 0)
        return 0;
    return batch_layout(n_gt, n_alleles_total, max_reps, monte_carlo ? 1 : 0).total;
}

static int max_slots_of(const SonicsParams *params, int n_gt, const int32_t *allele_off)
{
    int max_A = 0;
    for (int g = 0; g < n_gt; ++g)
        max_A = std::max(max_A, allele_off[g + 1] - allele_off[g]);
    return params->random_start ? max_A * max_A : max_A;
}

size_t sonics_genotype_work_bytes_for(const SonicsParams *params, int n_gt, const int32_t *allele_off, int max_reps)
{
    if (!params || !allele_off || n_gt <= 0 || max_reps <= 0)
        return 0;
    return batch_layout(n_gt, allele_off[n_gt], max_reps, params->monte_carlo ? 1 : 0,
                        max_slots_of(params, n_gt, allele_off)).total;
}

size_t sonics_genotype_work_bytes(int n_gt, int n_alleles_total, int max_reps)
{
    // sized for the larger (--monte_carlo) layout so one query serves both modes
    return sonics_genotype_work_bytes_mode(n_gt, n_alleles_total, max_reps, 1);
}

int sonics_genotype_batch(sonics_ctx *ctx, const SonicsParams *params, int n_gt, const int32_t *allele_off,
                          const int32_t *allele, const int32_t *reads, const float *noise_coef, const uint32_t *uid,
                          int max_reps, uint64_t seed, SonicsVerdict *verdict_out, void *dev_work, size_t dev_work_bytes,
                          void *stream)
{
    if (!ctx || !params || !allele_off || !allele || !reads || !verdict_out || !dev_work)
        return fail(SONICS_ERR_ARG, "sonics_genotype_batch: NULL argument");
    if (n_gt <= 0 || max_reps <= 0)
        return fail(SONICS_ERR_ARG, "sonics_genotype_batch: n_gt and max_reps must be positive");
    const int nA = allele_off[n_gt];
    const BatchLayout L = batch_layout(n_gt, nA, max_reps, params->monte_carlo, max_slots_of(params, n_gt, allele_off));
    if (dev_work_bytes < L.total)
        return fail(SONICS_ERR_SPACE, "sonics_genotype_batch: workspace %zu < %zu bytes", dev_work_bytes, L.total);
    unsigned char *w = (unsigned char *)dev_work;
    void *d_table = w + L.off_table;
    double *d_lnL = (double *)(w + L.off_lnL);
    uint16_t *d_group = (uint16_t *)(w + L.off_group);
    double *d_ident = params->monte_carlo ? (double *)(w + L.off_identity) : nullptr;
    float *d_rsq = params->monte_carlo ? (float *)(w + L.off_rsq) : nullptr;
    SonicsVerdict *d_verdict = (SonicsVerdict *)(w + L.off_verdict);
    int32_t *d_act[2] = {(int32_t *)(w + L.off_act0), (int32_t *)(w + L.off_act1)};
    int32_t *d_count = (int32_t *)(w + L.off_count);
    cudaStream_t s = (cudaStream_t)stream;

    if (int rc = sonics_table_build(ctx, params, n_gt, allele_off, allele, reads, noise_coef, uid, d_table,
                                    sonics_table_bytes(n_gt, nA), stream))
        return rc;
    struct TableGuard { // the table lives in the workspace: its host-side record goes away on every exit path
        sonics_ctx *c;
        const void *t;
        ~TableGuard() { c->tables.erase(t); }
    } table_guard{ctx, d_table};
    TableInfo ti;
    TableView tv;
    if (int rc = find_table(ctx, d_table, n_gt, ti, tv))
        return rc;
    CUDA_OK(cudaMemsetAsync(d_verdict, 0, (size_t)n_gt * sizeof(SonicsVerdict), s));

    // round schedule of sonics.pyx:79-82,193-199: cumulative 100, 500, 1000, 5000, ...
    int run = 0, rnd = 1;
    int reps = params->monte_carlo ? max_reps : (max_reps > 100 ? 100 : max_reps);
    const int32_t *active = nullptr; // all genotypes
    int n_active = n_gt, cur = 0;
    while (run < max_reps && n_active > 0) {
        if (int rc = sonics_simulate(ctx, params, d_table, n_gt, active, n_active, run, reps, seed, L.stride, run, d_lnL,
                                     d_group, d_ident, d_rsq, nullptr, w + L.off_scratch, L.scratch_bytes, stream))
            return rc;
        run += reps;
        const int final_round = run >= max_reps;
        if (int rc = launch_evaluate(ctx, params, ti, tv, active, n_active, run, L.stride, d_lnL, d_group, d_ident, d_rsq,
                                     final_round, d_verdict, w + L.off_scratch, L.scratch_bytes, stream))
            return rc;
        if (final_round || params->monte_carlo)
            break;
        // genotypes that passed leave the loop (:184-191); 4 bytes cross PCIe per round
        k_compact_active<<<1, 1024, 0, s>>>(active, n_active, d_verdict, d_act[cur], d_count);
        ctx->launches++;
        CUDA_OK(cudaGetLastError());
        int32_t h_count = 0;
        CUDA_OK(cudaMemcpyAsync(&h_count, d_count, 4, cudaMemcpyDeviceToHost, s));
        CUDA_OK(cudaStreamSynchronize(s));
        active = d_act[cur];
        cur ^= 1;
        n_active = h_count;
        reps = (rnd % 2 == 1) ? 4 * run : run;
        if (reps + run > max_reps)
            reps = max_reps - run;
        ++rnd;
    }
    if (!params->monte_carlo) {
        // K4: identity / r^2 of the reported row only
        if (int rc = sonics_replay_best(ctx, params, d_table, n_gt, nullptr, n_gt, seed, d_verdict, w + L.off_scratch,
                                        L.scratch_bytes, stream))
            return rc;
    }
    CUDA_OK(cudaMemcpyAsync(verdict_out, d_verdict, (size_t)n_gt * sizeof(SonicsVerdict), cudaMemcpyDeviceToHost, s));
    CUDA_OK(cudaStreamSynchronize(s));
    for (int g = 0; g < n_gt; ++g)
        if (verdict_out[g].best_sim == -2)
            return fail(SONICS_ERR_CUDA, "genotype %d: replay of the best simulation did not reproduce its lnL", g);
    return SONICS_OK;
}

} // extern "C"
