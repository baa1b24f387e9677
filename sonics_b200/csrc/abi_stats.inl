// extern "C" entry points for the statistics / repeat-until-pass path (included by sonics_b200.cu).

static int next_pow2(int v)
{
    int p = 1;
    while (p < v)
        p <<= 1;
    return p;
}

static int launch_evaluate(sonics_ctx *ctx, const SonicsParams *params, const TableInfo &ti, const TableView &tv,
                           const int32_t *dev_active, int n_active, int64_t n_run, int64_t out_stride,
                           const double *dev_lnL, const uint16_t *dev_group, const double *dev_identity,
                           const float *dev_rsq, int final_round, SonicsVerdict *dev_verdict, void *stream)
{
    if (n_run <= 0 || n_run > SONICS_STATS_MAX_N)
        return fail(SONICS_ERR_RANGE,
                    "statistics over %lld simulations per genotype: the on-chip path holds at most %d",
                    (long long)n_run, SONICS_STATS_MAX_N);
    if (params->n_add_ratios < 0 || params->n_add_ratios > SONICS_MAX_ADD_RATIOS)
        return fail(SONICS_ERR_ARG, "at most %d --add_ratios percentiles are supported", SONICS_MAX_ADD_RATIOS);
    const int n_slots = params->random_start ? ti.max_A * ti.max_A : ti.max_A;
    if (n_slots > SONICS_STATS_MAX_SLOTS)
        return fail(SONICS_ERR_RANGE, "%d genotype groups exceed the supported %d (use fewer alleles with --random)",
                    n_slots, SONICS_STATS_MAX_SLOTS);
    EvalArgs a;
    a.t = tv;
    a.active = dev_active;
    a.n_run = (int32_t)n_run;
    a.n_pad = std::max(next_pow2((int)n_run), 2);
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
    const size_t smem = stats_smem_bytes(a.n_pad, n_slots);
    CUDA_OK(cudaFuncSetAttribute(k_evaluate, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_evaluate<<<n_active, SONICS_STATS_THREADS, smem, (cudaStream_t)stream>>>(a, n_slots);
    ctx->launches++;
    CUDA_OK(cudaGetLastError());
    return SONICS_OK;
}

extern "C" {

int sonics_evaluate(sonics_ctx *ctx, const SonicsParams *params, const void *dev_table, int n_gt,
                    const int32_t *dev_active, int n_active, int64_t n_run, int64_t out_stride, const double *dev_lnL,
                    const uint16_t *dev_group, const double *dev_identity, const float *dev_rsq, int final_round,
                    SonicsVerdict *dev_verdict, void *stream)
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
                           dev_rsq, final_round, dev_verdict, stream);
}

int sonics_replay_best(sonics_ctx *ctx, const SonicsParams *params, const void *dev_table, int n_gt,
                       const int32_t *dev_active, int n_active, uint64_t seed, SonicsVerdict *dev_verdict, void *stream)
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
    a.n_active = n_active;
    a.sim_begin = 0;
    a.sim_count = 1;
    a.seed = seed;
    a.out_stride = 0;
    a.out_offset = 0;
    a.lnL = nullptr;
    a.group = nullptr;
    a.identity = nullptr;
    a.rsq = nullptr;
    a.simpar = nullptr;
    a.replay = dev_verdict;
    a.dbg = nullptr;
    return launch_simulate(ctx, a, ti, true, 1, stream);
}

// device workspace of sonics_genotype_batch
struct BatchLayout {
    size_t off_table, off_lnL, off_group, off_identity, off_rsq, off_verdict, off_act0, off_act1, off_count, total;
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
    L.total = o;
    return L;
}

size_t sonics_genotype_work_bytes(int n_gt, int n_alleles_total, int max_reps)
{
    if (n_gt <= 0 || n_alleles_total <= 0 || max_reps <= 0)
        return 0;
    // sized for the larger (--monte_carlo) layout so one query serves both modes
    return batch_layout(n_gt, n_alleles_total, max_reps, 1).total;
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
    if (max_reps > SONICS_STATS_MAX_N)
        return fail(SONICS_ERR_RANGE, "max_reps %d above the %d simulations per genotype the on-chip statistics hold",
                    max_reps, SONICS_STATS_MAX_N);
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
                                     d_group, d_ident, d_rsq, nullptr, stream))
            return rc;
        run += reps;
        const int final_round = run >= max_reps;
        if (int rc = launch_evaluate(ctx, params, ti, tv, active, n_active, run, L.stride, d_lnL, d_group, d_ident, d_rsq,
                                     final_round, d_verdict, stream))
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
        if (int rc = sonics_replay_best(ctx, params, d_table, n_gt, nullptr, n_gt, seed, d_verdict, stream))
            return rc;
    }
    CUDA_OK(cudaMemcpyAsync(verdict_out, d_verdict, (size_t)n_gt * sizeof(SonicsVerdict), cudaMemcpyDeviceToHost, s));
    CUDA_OK(cudaStreamSynchronize(s));
    for (int g = 0; g < n_gt; ++g)
        if (verdict_out[g].best_sim == -2)
            return fail(SONICS_ERR_CUDA, "genotype %d: replay of the best simulation did not reproduce its lnL", g);
    ctx->tables.erase(d_table);
    return SONICS_OK;
}

} // extern "C"
