// extern "C" entry points for the statistics / repeat-until-pass path (included by sonics_b200.cu).

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
}

struct SlotIsN { // indicator "sorted element i belongs to the target slot", 0 for the closing index n
    const uint16_t *slot;
    const BigState *st;
    int fixed; // >= 0: fixed target; -1: st->best; -2: st->second
    int n;
    __device__ uint32_t operator()(uint32_t i) const
    {
        if ((int)i >= n)
            return 0u;
        const int target = fixed >= 0 ? fixed : (fixed == -1 ? st->best : st->second);
        return slot[i] == target ? 1u : 0u;
    }
};

// Statistics of ONE genotype with n_run > SONICS_STATS_MAX_N simulations (see stats_big.cuh).
static int evaluate_big_one(sonics_ctx *ctx, const EvalArgs &a, const TableInfo &ti, int g, void *scratch,
                            size_t scratch_bytes, cudaStream_t s)
{
    const int n = a.n_run;
    const GtHeader &h = ti.hdr[g];
    const int A = h.n_alleles;
    const int n_slots = a.random_start ? A * A : A;
    if (scratch_bytes < stats_big_bytes(n, n_slots))
        return fail(SONICS_ERR_SPACE, "statistics scratch %zu < %zu bytes for %d simulations per genotype", scratch_bytes,
                    stats_big_bytes(n, n_slots), n);
    const int64_t base = (int64_t)g * a.out_stride;
    unsigned char *p = (unsigned char *)scratch;
    BigBufs b;
    auto take = [&](size_t bytes) {
        void *r = p;
        p += big_align(bytes);
        return r;
    };
    b.keys = (double *)take((size_t)n * 8);
    b.idx = (uint32_t *)take((size_t)n * 4);
    b.idx_in = (uint32_t *)take((size_t)n * 4);
    b.slot = (uint16_t *)take((size_t)n * 2);
    b.flag = (unsigned char *)take((size_t)n);
    b.heads = (uint32_t *)take((size_t)(n + 1) * 4);
    b.S_best = (uint32_t *)take((size_t)(n + 1) * 4);
    b.S = (uint32_t *)take((size_t)(n + 1) * 4);
    b.tmp_in = (double *)take((size_t)n * 8);
    b.tmp_out = (double *)take((size_t)n * 8);
    b.cnt = (int *)take((size_t)n_slots * 4);
    b.cnt_valid = (int *)take((size_t)n_slots * 4);
    b.maxk = (unsigned long long *)take((size_t)n_slots * 8);
    b.st = (BigState *)take(sizeof(BigState));
    b.cub_tmp = p;
    b.cub_bytes = scratch_bytes - (size_t)(p - (unsigned char *)scratch);
    const int T = 256, nb = (n + T - 1) / T, nbs = std::min(nb, ctx->sm_count * 8);
    int *n_heads_dev = &b.st->pad; // DeviceSelect writes the number of run heads here

    k_big_init<<<std::max(nb, (n_slots + T - 1) / T), T, 0, s>>>(b, n, n_slots);
    ctx->launches++;
    size_t need = 0;
    CUDA_OK(cub::DeviceRadixSort::SortPairs(nullptr, need, a.lnL + base, b.keys, b.idx_in, b.idx, n, 0, 64, s));
    if (need > b.cub_bytes)
        return fail(SONICS_ERR_SPACE, "CUB sort needs %zu bytes of temporary storage, %zu available", need, b.cub_bytes);
    need = b.cub_bytes;
    CUDA_OK(cub::DeviceRadixSort::SortPairs(b.cub_tmp, need, a.lnL + base, b.keys, b.idx_in, b.idx, n, 0, 64, s));
    ctx->launches++;
    k_big_slots<<<nb, T, 0, s>>>(b, a.group + base, n, A, h.first_idx, a.random_start);
    ctx->launches++;
    cub::CountingInputIterator<uint32_t> iota(0u);
    need = b.cub_bytes;
    CUDA_OK(cub::DeviceSelect::Flagged(b.cub_tmp, need, iota, b.flag, b.heads, n_heads_dev, n, s));
    ctx->launches++;
    k_big_top2<<<1, 1, 0, s>>>(b, n, n_slots, a.min_sim_opt, n_heads_dev);
    ctx->launches++;

    auto scan = [&](int fixed, uint32_t *out) -> int {
        SlotIsN f{b.slot, b.st, fixed, n};
        cub::TransformInputIterator<uint32_t, SlotIsN, cub::CountingInputIterator<uint32_t>> it(iota, f);
        size_t nb2 = b.cub_bytes;
        CUDA_OK(cub::DeviceScan::ExclusiveSum(b.cub_tmp, nb2, it, out, n + 1, s));
        ctx->launches++;
        return SONICS_OK;
    };

    if (a.monte_carlo) {
        k_big_best_row<<<1, 256, 0, s>>>(b, n, 1);
        k_big_set_best_slot<<<1, 1, 0, s>>>(b, a.group, base, A, h.first_idx, a.random_start);
        ctx->launches += 2;
        if (int rc = scan(-1, b.S_best))
            return rc;
        k_big_pick<<<nbs, T, 0, s>>>(b, b.S_best, n, 0, 0.75, 0);
        ctx->launches++;
        for (int which = 0; which < 2; ++which) {
            k_big_gather<<<nb, T, 0, s>>>(b, a.group, a.identity, a.rsq, base, n, A, h.first_idx, a.random_start, which);
            need = b.cub_bytes;
            CUDA_OK(cub::DeviceRadixSort::SortKeys(b.cub_tmp, need, b.tmp_in, b.tmp_out, n, 0, 64, s));
            k_big_q75_sorted<<<1, 1, 0, s>>>(b, which);
            ctx->launches += 3;
        }
        k_big_verdict<<<1, 1, 0, s>>>(b, a, g, n);
        ctx->launches++;
        CUDA_OK(cudaGetLastError());
        return SONICS_OK;
    }

    k_big_best_row<<<1, 256, 0, s>>>(b, n, 0);
    ctx->launches++;
    if (int rc = scan(-1, b.S_best))
        return rc;
    if (int rc = scan(-2, b.S))
        return rc;
    for (int qi = 0; qi <= a.n_add; ++qi) {
        const double q = qi == 0 ? 0.75 : a.add_q[qi - 1];
        k_big_pick<<<nbs, T, 0, s>>>(b, b.S_best, n, 0, q, 0);
        k_big_pick<<<nbs, T, 0, s>>>(b, b.S, n, 1, q, 2);
        k_big_ratio<<<1, 1, 0, s>>>(b, q, qi - 1);
        ctx->launches += 3;
    }
    for (int t = 0; t < n_slots; ++t) {
        k_big_pair_begin<<<1, 1, 0, s>>>(b);
        ctx->launches++;
        if (int rc = scan(t, b.S))
            return rc;
        k_big_pair<<<nbs, T, 0, s>>>(b, t);
        k_big_pair_end<<<1, 1, 0, s>>>(b, t);
        ctx->launches += 2;
    }
    k_big_verdict<<<1, 1, 0, s>>>(b, a, g, n);
    ctx->launches++;
    CUDA_OK(cudaGetLastError());
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
    if (n_slots > SONICS_STATS_MAX_SLOTS)
        return fail(SONICS_ERR_RANGE, "%d genotype groups exceed the supported %d (use fewer alleles with --random)",
                    n_slots, SONICS_STATS_MAX_SLOTS);
    EvalArgs a;
    fill_eval_args(a, params, tv, dev_active, n_run, out_stride, dev_lnL, dev_group, dev_identity, dev_rsq, final_round,
                   dev_verdict);
    cudaStream_t s = (cudaStream_t)stream;
    if (n_run <= SONICS_STATS_MAX_N) {
        const size_t smem = stats_smem_bytes(a.n_pad, n_slots);
        CUDA_OK(cudaFuncSetAttribute(k_evaluate, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_evaluate<<<n_active, SONICS_STATS_THREADS, smem, s>>>(a, n_slots);
        ctx->launches++;
        CUDA_OK(cudaGetLastError());
        return SONICS_OK;
    }
    // large-n path: one genotype at a time (typical use: a single locus with a large -r)
    if (!dev_scratch)
        return fail(SONICS_ERR_SPACE, "statistics over %lld simulations per genotype need a scratch buffer of "
                                      "sonics_stats_scratch_bytes() bytes",
                    (long long)n_run);
    std::vector<int32_t> act(n_active);
    if (dev_active) {
        CUDA_OK(cudaMemcpyAsync(act.data(), dev_active, (size_t)n_active * 4, cudaMemcpyDeviceToHost, s));
        CUDA_OK(cudaStreamSynchronize(s));
    } else {
        for (int i = 0; i < n_active; ++i)
            act[i] = i;
    }
    for (int i = 0; i < n_active; ++i)
        if (int rc = evaluate_big_one(ctx, a, ti, act[i], dev_scratch, dev_scratch_bytes, s))
            return rc;
    return SONICS_OK;
}

extern "C" {

size_t sonics_stats_scratch_bytes(int64_t n_run, int max_alleles, int random_start)
{
    if (n_run <= SONICS_STATS_MAX_N)
        return 0;
    return stats_big_bytes(n_run, random_start ? max_alleles * max_alleles : max_alleles);
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

static BatchLayout batch_layout(int n_gt, int n_alleles_total, int max_reps, int monte_carlo)
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
    if (max_reps > SONICS_STATS_MAX_N) // the large-n statistics reuse the scratch between rounds
        L.scratch_bytes = std::max(L.scratch_bytes, stats_big_bytes(max_reps, SONICS_STATS_MAX_SLOTS));
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
    const BatchLayout L = batch_layout(n_gt, nA, max_reps, params->monte_carlo);
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
