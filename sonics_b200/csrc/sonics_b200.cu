// C-ABI of the B200-native SONiCS engine (see include/sonics_b200.h).
// Host side: argument checking, genotype-table packing, launch configuration.
// No device memory is allocated here; all device buffers belong to the caller.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include <cub/cub.cuh>

#include "../../include/sonics_b200.h"
#include "sim_kernel.cuh"
#include "sim_lockstep.cuh"
#include "stats_kernel.cuh"

using namespace sonics;

static thread_local std::string g_err;

static int fail(int code, const char *fmt, ...)
{
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}

#define CUDA_OK(expr)                                                                                        \
    do {                                                                                                     \
        cudaError_t e__ = (expr);                                                                            \
        if (e__ != cudaSuccess)                                                                              \
            return fail(SONICS_ERR_CUDA, "%s failed: %s", #expr, cudaGetErrorString(e__));                   \
    } while (0)

struct TableInfo {
    int n_gt, n_alleles, max_W, max_A;
    int min_al, max_al, max_span, any_noise; // lowest / highest input allele, widest input range of a genotype,
                                             // noise molecules anywhere (lockstep sort key)
    size_t off_allele, off_reads, off_fl;
    std::vector<GtHeader> hdr; // host copy of the headers (the large-n statistics path needs A / first_idx)
};

struct sonics_ctx {
    int device;
    int sm_count;
    int64_t launches;
    int max_W_last;
    unsigned long long *dbg;
    SonicsTraceEntry *trace; // sonics_debug_trace
    int32_t *trace_len;
    int32_t trace_cap;
    int32_t *readout_dump;   // sonics_debug_readout
    int kernel;              // SONICS_KERNEL_*: which form of K1 sonics_simulate launches
    int sort_bits[3];        // lockstep sort key: bits of efficiency / down / up (-1 = automatic)
    int sort_mode;           // SONICS_SORT_*
    float fast_npq;          // > 0: approximate sampler for binomial draws with n p q >= this (sonics_set_samplers)
    std::map<const void *, TableInfo> tables;
};

static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

static void table_layout(int n_gt, int n_alleles, TableInfo &ti)
{
    ti.n_gt = n_gt;
    ti.n_alleles = n_alleles;
    ti.off_allele = align_up((size_t)n_gt * sizeof(GtHeader), 16);
    ti.off_reads = ti.off_allele + align_up((size_t)n_alleles * 4, 16);
    ti.off_fl = ti.off_reads + align_up((size_t)n_alleles * 4, 16);
}

extern "C" {

int sonics_abi_version(void) { return SONICS_ABI_VERSION; }
const char *sonics_last_error(void) { return g_err.c_str(); }

int sonics_create(int device, sonics_ctx **out)
{
    if (!out)
        return fail(SONICS_ERR_ARG, "sonics_create: out is NULL");
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0)
        return fail(SONICS_ERR_CUDA, "sonics_create: no CUDA device (%s); this library has no CPU fallback",
                    e == cudaSuccess ? "count = 0" : cudaGetErrorString(e));
    if (device < 0 || device >= n)
        return fail(SONICS_ERR_ARG, "sonics_create: device %d out of range (%d devices)", device, n);
    cudaDeviceProp prop;
    CUDA_OK(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10)
        return fail(SONICS_ERR_CUDA, "sonics_create: device %d is sm_%d%d; this build carries sm_100a code only", device,
                    prop.major, prop.minor);
    CUDA_OK(cudaSetDevice(device));
    sonics_ctx *c = new sonics_ctx();
    c->device = device;
    c->sm_count = prop.multiProcessorCount;
    c->launches = 0;
    c->max_W_last = 0;
    c->dbg = nullptr;
    c->trace = nullptr;
    c->trace_len = nullptr;
    c->trace_cap = 0;
    c->readout_dump = nullptr;
    c->kernel = SONICS_KERNEL_AUTO;
    c->sort_bits[0] = c->sort_bits[1] = c->sort_bits[2] = -1;
    if (const char *e = getenv("SONICS_K1")) { // measurement switch; tests use sonics_set_kernel
        if (!strcmp(e, "persistent"))
            c->kernel = SONICS_KERNEL_PERSISTENT;
        else if (!strcmp(e, "lockstep"))
            c->kernel = SONICS_KERNEL_LOCKSTEP;
    }
    if (const char *e = getenv("SONICS_SORT_BITS"))
        sscanf(e, "%d,%d,%d", &c->sort_bits[0], &c->sort_bits[1], &c->sort_bits[2]);
    c->sort_mode = SONICS_SORT_PAIR_MAJOR;
    c->fast_npq = 0.0f;
    if (const char *e = getenv("SONICS_SORT_MODE"))
        if (!strcmp(e, "slot"))
            c->sort_mode = SONICS_SORT_SLOT_MAJOR;
    *out = c;
    return SONICS_OK;
}

void sonics_destroy(sonics_ctx *ctx) { delete ctx; }

int sonics_sync(sonics_ctx *ctx, void *stream)
{
    if (!ctx)
        return fail(SONICS_ERR_ARG, "sonics_sync: ctx is NULL");
    CUDA_OK(cudaSetDevice(ctx->device));
    CUDA_OK(cudaStreamSynchronize((cudaStream_t)stream));
    return SONICS_OK;
}

int64_t sonics_launch_count(sonics_ctx *ctx) { return ctx ? ctx->launches : 0; }

int sonics_mvhyperg(sonics_ctx *ctx, const int64_t *x, const int64_t *color, int k, int n_pools, double *like_out,
                    void *dev_scratch, size_t dev_scratch_bytes)
{
    if (!ctx || !x || !color || !like_out || !dev_scratch)
        return fail(SONICS_ERR_ARG, "sonics_mvhyperg: NULL argument");
    if (k <= 0 || n_pools <= 0)
        return fail(SONICS_ERR_ARG, "sonics_mvhyperg: k and n_pools must be positive");
    const size_t row = (size_t)k * 8, need = (size_t)n_pools * (2 * row + 8);
    if (dev_scratch_bytes < need)
        return fail(SONICS_ERR_SPACE, "sonics_mvhyperg: scratch %zu < %zu bytes", dev_scratch_bytes, need);
    CUDA_OK(cudaSetDevice(ctx->device));
    int64_t *dx = (int64_t *)dev_scratch;
    int64_t *dc = dx + (size_t)n_pools * k;
    double *dout = (double *)(dc + (size_t)n_pools * k);
    CUDA_OK(cudaMemcpy(dx, x, (size_t)n_pools * row, cudaMemcpyHostToDevice));
    CUDA_OK(cudaMemcpy(dc, color, (size_t)n_pools * row, cudaMemcpyHostToDevice));
    k_mvhyperg<<<(n_pools + 127) / 128, 128>>>(dx, dc, k, n_pools, dout);
    ctx->launches++;
    CUDA_OK(cudaGetLastError());
    CUDA_OK(cudaMemcpy(like_out, dout, (size_t)n_pools * 8, cudaMemcpyDeviceToHost));
    return SONICS_OK;
}

size_t sonics_table_bytes(int n_gt, int n_alleles_total)
{
    if (n_gt < 0 || n_alleles_total < 0)
        return 0;
    TableInfo ti;
    table_layout(n_gt, n_alleles_total, ti);
    return ti.off_fl + align_up((size_t)n_alleles_total * 8, 16);
}

static int check_params(const SonicsParams *p)
{
    if (!p)
        return fail(SONICS_ERR_ARG, "params is NULL");
    if (p->n_cycles < 0 || p->n_cycles > 400)
        return fail(SONICS_ERR_ARG, "n_cycles %d outside [0, 400]", p->n_cycles);
    if (p->start_copies < 0)
        return fail(SONICS_ERR_ARG, "start_copies must be >= 0");
    // counts are held in int32 on the device
    const double growth = (double)p->start_copies * std::pow(1.0 + std::max(0.0, p->efficiency[1]), p->n_cycles);
    if (growth * 1.25 >= 2147483647.0)
        return fail(SONICS_ERR_RANGE, "start_copies * (1 + efficiency_max)^n_cycles = %.3g overflows int32 counts",
                    growth);
    return SONICS_OK;
}

static SimParamsDev dev_params(const SonicsParams *p)
{
    const double small_number = 1e-16; // sonics.pyx:36
    SimParamsDev d;
    d.n_cycles = p->n_cycles;
    d.capture_cycle = p->capture_cycle;
    d.random_start = p->random_start;
    d.up_preference = p->up_preference;
    d.half_copies = (int32_t)((double)p->start_copies / 2.0); // :326-327, truncation into the int array
    d.down_lo = p->down[0] + small_number;
    d.down_rng = p->down[1] - d.down_lo;
    d.up_lo = p->up[0] + small_number;
    d.up_hi = p->up[1];
    d.cap_lo = p->capture[0] + small_number;
    d.cap_rng = p->capture[1] - d.cap_lo;
    d.eff_lo = p->efficiency[0] + small_number;
    d.eff_rng = p->efficiency[1] - d.eff_lo;
    return d;
}

int sonics_table_build(sonics_ctx *ctx, const SonicsParams *params, int n_gt, const int32_t *allele_off,
                       const int32_t *allele, const int32_t *reads, const float *noise_coef, const uint32_t *uid,
                       void *dev_table, size_t dev_table_bytes, void *stream)
{
    if (!ctx || !allele_off || !allele || !reads || !dev_table)
        return fail(SONICS_ERR_ARG, "sonics_table_build: NULL argument");
    if (int rc = check_params(params))
        return rc;
    if (n_gt <= 0)
        return fail(SONICS_ERR_ARG, "sonics_table_build: n_gt must be positive");
    const int nA = allele_off[n_gt];
    if (allele_off[0] != 0 || nA <= 0)
        return fail(SONICS_ERR_ARG, "sonics_table_build: malformed allele_off");
    const size_t need = sonics_table_bytes(n_gt, nA);
    if (dev_table_bytes < need)
        return fail(SONICS_ERR_SPACE, "sonics_table_build: table buffer %zu < %zu bytes", dev_table_bytes, need);
    TableInfo ti;
    table_layout(n_gt, nA, ti);
    std::vector<unsigned char> host(ti.off_fl, 0);
    GtHeader *hdr = (GtHeader *)host.data();
    int32_t *h_allele = (int32_t *)(host.data() + ti.off_allele);
    int32_t *h_reads = (int32_t *)(host.data() + ti.off_reads);
    int max_W = 0, max_A = 0, min_al = SONICS_MAX_ALLELE, max_al = 0, max_span = 0, any_noise = 0;
    const int64_t sum0 = 2 * (int64_t)((double)params->start_copies / 2.0);
    for (int g = 0; g < n_gt; ++g) {
        const int a0 = allele_off[g], a1 = allele_off[g + 1], A = a1 - a0;
        if (A < 2)
            return fail(SONICS_ERR_ALLELES, "genotype %d: Less then two alleles as starting conditions! Aborting.", g);
        if (A > 255)
            return fail(SONICS_ERR_RANGE, "genotype %d: %d alleles (max 255)", g, A);
        int64_t D = 0;
        int first = 0;
        for (int i = a0; i < a1; ++i) {
            if (allele[i] <= 0 || allele[i] >= SONICS_MAX_ALLELE)
                return fail(SONICS_ERR_RANGE, "genotype %d: allele %d outside (0, %d)", g, allele[i], SONICS_MAX_ALLELE);
            if (i > a0 && allele[i] <= allele[i - 1])
                return fail(SONICS_ERR_ARG, "genotype %d: alleles must be strictly ascending", g);
            if (reads[i] <= 0)
                return fail(SONICS_ERR_ARG, "genotype %d: reads must be positive", g);
            D += reads[i];
            if (reads[i] > reads[a0 + first])
                first = i - a0; // np.argmax: first occurrence of the maximum (:323)
            h_allele[i] = allele[i];
            h_reads[i] = reads[i];
        }
        if (D >= 2147483647)
            return fail(SONICS_ERR_RANGE, "genotype %d: read total overflows int32", g);
        const int min_in = allele[a0], max_in = allele[a1 - 1];
        int floor_al;
        if (params->floor_opt == -1)
            floor_al = 1;
        else
            floor_al = std::max(1, min_in - params->floor_opt);
        // bins that can ever hold molecules: down-stutter walks one bin per cycle and stops at
        // floor-1 (:427,472); up-stutter walks one bin per cycle (:476)
        int lo = std::max(floor_al - 1, min_in - params->n_cycles);
        lo = std::max(0, std::min(lo, min_in));
        // The reference's histogram ends at bin SONICS_MAX_ALLELE - 1 (:21); up-stutter out of that bin raises an
        // IndexError there (:476).  Here the window is clamped to it and such molecules are dropped: a genotype
        // near the end of the range is simulated like any other instead of failing the whole batch.
        const int hi = std::min(max_in + params->n_cycles, SONICS_MAX_ALLELE - 1);
        GtHeader &h = hdr[g];
        h.a_off = a0;
        h.n_alleles = A;
        h.lo = lo;
        h.W = hi - lo + 1;
        h.floor_al = floor_al;
        h.first_idx = first;
        h.D = (int32_t)D;
        const float nc = noise_coef ? noise_coef[g] : 0.0f;
        h.noise_add = nc > 0.0f ? (int32_t)((double)nc * (double)sum0) : 0; // :334-337
        h.uid = uid ? uid[g] : (uint32_t)g;
        h.in_lo_bin = min_in - lo;
        h.fl_D = 0.0;
        max_W = std::max(max_W, h.W);
        max_A = std::max(max_A, A);
        min_al = std::min(min_al, min_in);
        max_al = std::max(max_al, max_in);
        max_span = std::max(max_span, max_in - min_in);
        any_noise |= h.noise_add > 0;
    }
    ti.max_W = max_W;
    ti.max_A = max_A;
    ti.min_al = min_al;
    ti.max_al = max_al;
    ti.max_span = max_span;
    ti.any_noise = any_noise;
    ti.hdr.assign(hdr, hdr + n_gt);
    CUDA_OK(cudaSetDevice(ctx->device));
    cudaStream_t s = (cudaStream_t)stream;
    CUDA_OK(cudaMemcpyAsync(dev_table, host.data(), host.size(), cudaMemcpyHostToDevice, s));
    // pageable source: the copy is staged before the call returns, `host` may go out of scope
    unsigned char *base = (unsigned char *)dev_table;
    const int n_thr = std::max(n_gt, nA);
    k_table_prep<<<(n_thr + 255) / 256, 256, 0, s>>>((GtHeader *)base, (const int32_t *)(base + ti.off_reads),
                                                      (double *)(base + ti.off_fl), n_gt, nA);
    ctx->launches++;
    CUDA_OK(cudaGetLastError());
    ctx->tables[dev_table] = ti;
    ctx->max_W_last = max_W;
    return SONICS_OK;
}

int sonics_table_max_window(sonics_ctx *ctx) { return ctx ? ctx->max_W_last : 0; }

int sonics_table_release(sonics_ctx *ctx, const void *dev_table)
{
    if (!ctx)
        return fail(SONICS_ERR_ARG, "sonics_table_release: ctx is NULL");
    ctx->tables.erase(dev_table);
    return SONICS_OK;
}

static int find_table(sonics_ctx *ctx, const void *dev_table, int n_gt, TableInfo &ti, TableView &tv)
{
    auto it = ctx->tables.find(dev_table);
    if (it == ctx->tables.end())
        return fail(SONICS_ERR_ARG, "unknown genotype table %p (call sonics_table_build first)", dev_table);
    ti = it->second;
    if (ti.n_gt != n_gt)
        return fail(SONICS_ERR_ARG, "table holds %d genotypes, call says %d", ti.n_gt, n_gt);
    const unsigned char *base = (const unsigned char *)dev_table;
    tv.hdr = (const GtHeader *)base;
    tv.allele = (const int32_t *)(base + ti.off_allele);
    tv.reads = (const int32_t *)(base + ti.off_reads);
    tv.fl_reads = (const double *)(base + ti.off_fl);
    return SONICS_OK;
}

// Scratch of the persistent form: [counters: 256 B][pool: ceil(units / 32) * Wcap * 32 int32]
static size_t scratch_bytes_for(int64_t n_units, int Wcap)
{
    const int64_t tiles = (n_units + 31) / 32;
    return 256 + (size_t)tiles * Wcap * 32 * 4;
}

// Scratch of the lockstep form: [counters: 256 B][keys x2: u64][units x2: u32][CUB temporary storage]
struct LsLayout {
    size_t off_k0, off_k1, off_v0, off_v1, off_tmp, tmp_bytes, total;
};
static int ls_layout(int64_t n, LsLayout &L)
{
    size_t o = 256;
    L.off_k0 = o;
    o += align_up((size_t)n * 8, 256);
    L.off_k1 = o;
    o += align_up((size_t)n * 8, 256);
    L.off_v0 = o;
    o += align_up((size_t)n * 4, 256);
    L.off_v1 = o;
    o += align_up((size_t)n * 4, 256);
    L.off_tmp = o;
    cub::DoubleBuffer<uint64_t> dk(nullptr, nullptr);
    cub::DoubleBuffer<uint32_t> dv(nullptr, nullptr);
    size_t tmp = 0;
    if (cub::DeviceRadixSort::SortPairs(nullptr, tmp, dk, dv, (int)n, 0, 64) != cudaSuccess)
        return -1;
    L.tmp_bytes = align_up(tmp, 256) + 256;
    L.total = o + L.tmp_bytes;
    return 0;
}

static int bits_for(int64_t v) // bits that hold the values 0 .. v-1
{
    int b = 0;
    while (((int64_t)1 << b) < v)
        ++b;
    return b;
}

// Persistent form (K1a k_pcr + K1b k_lnl): units [0, total_units) of `a` (unit -> (slot, jrel) as documented at
// SimArgs) in chunks that fit the caller's scratch: per chunk one persistent k_pcr launch and one k_lnl launch.
static int launch_simulate_persistent(sonics_ctx *ctx, SimArgs &a, const TableInfo &ti, bool readout,
                                      int64_t total_units, void *dev_scratch, size_t scratch_bytes, void *stream)
{
    cudaStream_t s = (cudaStream_t)stream;
    a.Wcap = ti.max_W;
    a.dbg = ctx->dbg;
    a.trace = ctx->trace;
    a.trace_len = ctx->trace_len;
    a.trace_cap = ctx->trace_cap;
    a.readout_dump = ctx->readout_dump;
    a.order = nullptr;
    auto pcr = (ctx->trace || ctx->dbg) ? k_pcr<true> : k_pcr<false>;
    if (!dev_scratch || scratch_bytes < scratch_bytes_for(32, a.Wcap))
        return fail(SONICS_ERR_SPACE, "simulation scratch of %zu bytes is below the minimum %zu", scratch_bytes,
                    scratch_bytes_for(32, a.Wcap));
    const int64_t cap_units = std::min<int64_t>(((int64_t)(scratch_bytes - 256) / ((int64_t)a.Wcap * 128)) * 32,
                                                 (int64_t)1 << 30);
    a.counter = (unsigned int *)dev_scratch;
    a.pool = (int32_t *)((unsigned char *)dev_scratch + 256);
    const size_t per_warp = warp_smem_bytes(a.Wcap);
    int wpb = (int)std::min<size_t>(4, (200 * 1024) / per_warp);
    if (wpb < 1)
        return fail(SONICS_ERR_RANGE, "window of %d bins does not fit shared memory", a.Wcap);
    const size_t smem = SONICS_RCP_TABLE * 4 + per_warp * wpb;
    CUDA_OK(cudaFuncSetAttribute(pcr, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int occ = 0;
    CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, pcr, wpb * 32, smem));
    if (occ < 1)
        occ = 1;
    auto lnl = readout ? k_lnl<true> : k_lnl<false>;
    if (ctx->trace && total_units > cap_units)
        return fail(SONICS_ERR_SPACE, "draw trace: %lld simulations do not fit one scratch chunk of %lld",
                    (long long)total_units, (long long)cap_units);
    for (int64_t u0 = 0; u0 < total_units; u0 += cap_units) {
        const int64_t nu = std::min<int64_t>(cap_units, total_units - u0);
        a.unit_base = u0;
        a.n_units = (int32_t)nu;
        CUDA_OK(cudaMemsetAsync(a.counter, 0, 256, s));
        // persistent lanes: no more blocks than there are lanes' worth of work
        const int64_t want = (nu + wpb * 32 - 1) / (wpb * 32);
        const int grid = (int)std::min<int64_t>(want, (int64_t)ctx->sm_count * occ);
        pcr<<<grid, wpb * 32, smem, s>>>(a);
        ctx->launches++;
        CUDA_OK(cudaGetLastError());
        const int lgrid = (int)std::min<int64_t>((nu + 127) / 128, (int64_t)ctx->sm_count * 16);
        lnl<<<lgrid, 128, 0, s>>>(a);
        ctx->launches++;
        CUDA_OK(cudaGetLastError());
    }
    return SONICS_OK;
}

// Lockstep form (sim_lockstep.cuh): per chunk k_sort_keys + CUB radix sort + k_pcr_ls (lnL / readout fused).
static int launch_simulate_lockstep(sonics_ctx *ctx, SimArgs &a, const TableInfo &ti, bool readout, int64_t total_units,
                                    int n_slots, void *dev_scratch, size_t scratch_bytes, void *stream)
{
    cudaStream_t s = (cudaStream_t)stream;
    a.Wcap = ti.max_W;
    a.dbg = nullptr;
    a.trace = nullptr;
    a.trace_len = nullptr;
    a.trace_cap = 0;
    a.readout_dump = ctx->readout_dump;
    a.pool = nullptr;
    // chunk size from the scratch
    LsLayout L;
    if (ls_layout(32, L))
        return fail(SONICS_ERR_CUDA, "CUB temporary-storage query failed");
    if (!dev_scratch || scratch_bytes < L.total)
        return fail(SONICS_ERR_SPACE, "simulation scratch of %zu bytes is below the minimum %zu", scratch_bytes, L.total);
    int64_t cap_units = std::min<int64_t>(total_units, (int64_t)1 << 30);
    while (true) {
        if (ls_layout(cap_units, L))
            return fail(SONICS_ERR_CUDA, "CUB temporary-storage query failed");
        if (L.total <= scratch_bytes)
            break;
        cap_units = std::max<int64_t>(32, (int64_t)((double)cap_units * (double)scratch_bytes / (double)L.total * 0.98));
    }
    // sort key: slot | group | efficiency | down | up
    const int A = ti.max_A;
    const int64_t n_groups = a.p.random_start ? (int64_t)A * A : A;
    SortCfg sc;
    sc.grp_bits = bits_for(n_groups);
    const int slot_bits = bits_for(n_slots);
    // pool the simulations of equal start pairs across the genotypes of the launch (see SortCfg)
    sc.pair_major = (a.sims_per_gt > 1 && ctx->sort_mode != SONICS_SORT_SLOT_MAJOR) ? 1 : 0;
    sc.al_base = ti.min_al;
    sc.al_bits = bits_for((int64_t)ti.max_al - ti.min_al + 1);
    sc.span_bits = bits_for((int64_t)ti.max_span + 1);
    sc.noise_bits = ti.any_noise ? 1 : 0;
    // one genotype per launch: a bit more per prior (+0.5 ... 1 %); with many genotypes finer fields scatter the
    // lanes of a warp over the table and the result rows (-4 % on the cfg4 loop)
    sc.morton_bits = SONICS_LS_MORTON_BITS + (n_slots == 1 ? 1 : 0);
    const int pair_bits = sc.noise_bits + sc.al_bits + sc.span_bits * (sc.noise_bits ? 3 : 1);
    // 32-bit keys (8 instead of 12 bytes per unit and pass) when one bit less per prior makes them fit
    if (pair_bits + 3 * sc.morton_bits > 32 && pair_bits + 3 * (sc.morton_bits - 1) <= 32)
        sc.morton_bits -= 1;
    const double per_group = (double)a.sims_per_gt / (double)(a.p.random_start ? (int64_t)A * (A + 1) / 2 : A);
    int pbits = 0;
    while (pbits < 15 && per_group / (double)(1 << (pbits + 1)) >= 32.0)
        ++pbits;
    sc.eff_bits = (pbits + 1) / 2;
    sc.dn_bits = (pbits - sc.eff_bits + 1) / 2;
    sc.up_bits = pbits - sc.eff_bits - sc.dn_bits;
    if (ctx->sort_bits[0] >= 0) {
        sc.eff_bits = std::min(ctx->sort_bits[0], 15);
        sc.dn_bits = std::min(std::max(ctx->sort_bits[1], 0), 15);
        sc.up_bits = std::min(std::max(ctx->sort_bits[2], 0), 15);
    }
    // one simulation per genotype (K4 replay): the list is already in genotype order
    const int key_bits = a.sims_per_gt == 1 ? 0
                         : sc.pair_major ? pair_bits + 3 * sc.morton_bits
                                         : slot_bits + sc.grp_bits + sc.eff_bits + sc.dn_bits + sc.up_bits;
    if (key_bits > 64)
        return fail(SONICS_ERR_RANGE, "sort key of %d bits", key_bits);

    a.fast_npq = ctx->fast_npq;
    auto pcr = ctx->fast_npq > 0.0f ? (readout ? k_pcr_ls<true, true> : k_pcr_ls<false, true>)
                                    : (readout ? k_pcr_ls<true, false> : k_pcr_ls<false, false>);
    const size_t per_warp = warp_smem_bytes(a.Wcap);
    int wpb = (int)std::min<size_t>(4, (200 * 1024) / per_warp);
    if (wpb < 1)
        return fail(SONICS_ERR_RANGE, "window of %d bins does not fit shared memory", a.Wcap);
    const size_t smem = SONICS_RCP_TABLE * 4 + per_warp * wpb;
    CUDA_OK(cudaFuncSetAttribute(pcr, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int occ = 0;
    CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, pcr, wpb * 32, smem));
    if (occ < 1)
        occ = 1;
    unsigned char *base = (unsigned char *)dev_scratch;
    a.counter = (unsigned int *)base;
    for (int64_t u0 = 0; u0 < total_units; u0 += cap_units) {
        const int64_t nu = std::min<int64_t>(cap_units, total_units - u0);
        a.unit_base = u0;
        a.n_units = (int32_t)nu;
        CUDA_OK(cudaMemsetAsync(a.counter, 0, 256, s));
        const int kgrid = (int)((nu + 255) / 256);
        uint32_t *v0 = (uint32_t *)(base + L.off_v0), *v1 = (uint32_t *)(base + L.off_v1);
        size_t tmp = L.tmp_bytes;
        if (key_bits == 0) {
            a.order = nullptr;
        } else if (key_bits <= 32) {
            uint32_t *k0 = (uint32_t *)(base + L.off_k0), *k1 = (uint32_t *)(base + L.off_k1);
            k_sort_keys<uint32_t><<<kgrid, 256, 0, s>>>(a, sc, k0, v0);
            cub::DoubleBuffer<uint32_t> dk(k0, k1);
            cub::DoubleBuffer<uint32_t> dv(v0, v1);
            CUDA_OK(cub::DeviceRadixSort::SortPairs(base + L.off_tmp, tmp, dk, dv, (int)nu, 0, key_bits, s));
            a.order = dv.Current();
            ctx->launches += 2;
        } else {
            uint64_t *k0 = (uint64_t *)(base + L.off_k0), *k1 = (uint64_t *)(base + L.off_k1);
            k_sort_keys<uint64_t><<<kgrid, 256, 0, s>>>(a, sc, k0, v0);
            cub::DoubleBuffer<uint64_t> dk(k0, k1);
            cub::DoubleBuffer<uint32_t> dv(v0, v1);
            CUDA_OK(cub::DeviceRadixSort::SortPairs(base + L.off_tmp, tmp, dk, dv, (int)nu, 0, key_bits, s));
            a.order = dv.Current();
            ctx->launches += 2;
        }
        CUDA_OK(cudaGetLastError());
        const int64_t batches = (nu + 31) / 32;
        const int64_t want = (batches + wpb - 1) / wpb;
        const int grid = (int)std::min<int64_t>(want, (int64_t)ctx->sm_count * occ);
        pcr<<<grid, wpb * 32, smem, s>>>(a);
        ctx->launches++;
        CUDA_OK(cudaGetLastError());
    }
    return SONICS_OK;
}

static int launch_simulate(sonics_ctx *ctx, SimArgs &a, const TableInfo &ti, bool readout, int64_t total_units,
                           int n_slots, void *dev_scratch, size_t scratch_bytes, void *stream)
{
    // the draw trace and the control-flow counters exist in the persistent form only
    const bool persistent = ctx->kernel == SONICS_KERNEL_PERSISTENT || ctx->trace || ctx->dbg;
    if (persistent)
        return launch_simulate_persistent(ctx, a, ti, readout, total_units, dev_scratch, scratch_bytes, stream);
    return launch_simulate_lockstep(ctx, a, ti, readout, total_units, n_slots, dev_scratch, scratch_bytes, stream);
}

size_t sonics_scratch_bytes_kernel(int64_t n_units, int max_window, int kernel)
{
    if (n_units <= 0 || max_window <= 0)
        return 0;
    if (kernel == SONICS_KERNEL_PERSISTENT)
        return scratch_bytes_for(n_units, max_window);
    LsLayout L;
    if (ls_layout(n_units, L))
        return 0;
    return L.total;
}

size_t sonics_scratch_bytes(int64_t n_units, int max_window)
{
    return sonics_scratch_bytes_kernel(n_units, max_window, SONICS_KERNEL_LOCKSTEP);
}

#ifdef SONICS_LS_STATS
// diagnostics build only: read and clear the lockstep schedule counters (tools/ls_stats.py)
int sonics_debug_ls_stats(uint64_t *host_out128)
{
    unsigned long long z[128] = {0};
    if (cudaMemcpyFromSymbol(host_out128, g_ls_stats, sizeof(z)) != cudaSuccess)
        return SONICS_ERR_CUDA;
    if (cudaMemcpyToSymbol(g_ls_stats, z, sizeof(z)) != cudaSuccess)
        return SONICS_ERR_CUDA;
    return SONICS_OK;
}
#endif

int sonics_set_kernel(sonics_ctx *ctx, int kernel)
{
    if (!ctx || kernel < SONICS_KERNEL_AUTO || kernel > SONICS_KERNEL_LOCKSTEP)
        return fail(SONICS_ERR_ARG, "sonics_set_kernel: bad argument");
    ctx->kernel = kernel;
    return SONICS_OK;
}

int sonics_set_samplers(sonics_ctx *ctx, int mode, double min_npq)
{
    if (!ctx || (mode != SONICS_SAMPLERS_EXACT && mode != SONICS_SAMPLERS_FAST))
        return fail(SONICS_ERR_ARG, "sonics_set_samplers: bad argument");
    if (mode == SONICS_SAMPLERS_FAST && !(min_npq >= 20.0))
        return fail(SONICS_ERR_RANGE, "sonics_set_samplers: the approximate sampler needs n p q >= 20 (got %g)", min_npq);
    ctx->fast_npq = mode == SONICS_SAMPLERS_FAST ? (float)min_npq : 0.0f;
    return SONICS_OK;
}

int sonics_set_sort_mode(sonics_ctx *ctx, int mode)
{
    if (!ctx || (mode != SONICS_SORT_PAIR_MAJOR && mode != SONICS_SORT_SLOT_MAJOR))
        return fail(SONICS_ERR_ARG, "sonics_set_sort_mode: bad argument");
    ctx->sort_mode = mode;
    return SONICS_OK;
}

int sonics_set_sort_bits(sonics_ctx *ctx, int eff_bits, int down_bits, int up_bits)
{
    if (!ctx || eff_bits > 15 || down_bits > 15 || up_bits > 15)
        return fail(SONICS_ERR_ARG, "sonics_set_sort_bits: bad argument");
    ctx->sort_bits[0] = eff_bits;
    ctx->sort_bits[1] = down_bits;
    ctx->sort_bits[2] = up_bits;
    return SONICS_OK;
}

int sonics_debug_readout(sonics_ctx *ctx, int32_t *dev_readout)
{
    if (!ctx)
        return fail(SONICS_ERR_ARG, "sonics_debug_readout: ctx is NULL");
    ctx->readout_dump = dev_readout;
    return SONICS_OK;
}

int sonics_simulate(sonics_ctx *ctx, const SonicsParams *params, const void *dev_table, int n_gt,
                    const int32_t *dev_active, int n_active, int64_t sim_begin, int64_t sim_count, uint64_t seed,
                    int64_t out_stride, int64_t out_offset, double *dev_lnL, uint16_t *dev_group,
                    double *dev_identity, float *dev_rsq, double *dev_simpar, void *dev_scratch,
                    size_t dev_scratch_bytes, void *stream)
{
    if (!ctx || !dev_table || !dev_lnL || !dev_group)
        return fail(SONICS_ERR_ARG, "sonics_simulate: NULL argument");
    if (int rc = check_params(params))
        return rc;
    if (sim_count <= 0 || sim_begin < 0)
        return fail(SONICS_ERR_ARG, "sonics_simulate: bad simulation range");
    if (sim_count > 2147483647LL)
        return fail(SONICS_ERR_RANGE, "sonics_simulate: sim_count too large");
    if (out_offset < 0 || out_stride < out_offset + sim_count)
        return fail(SONICS_ERR_SPACE, "sonics_simulate: out_stride %lld < out_offset + sim_count",
                    (long long)out_stride);
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
    a.sims_per_gt = (int32_t)sim_count;
    a.sim_begin = sim_begin;
    a.seed = seed;
    a.out_stride = out_stride;
    a.out_offset = out_offset;
    a.lnL = dev_lnL;
    a.group = dev_group;
    a.identity = dev_identity;
    a.rsq = dev_rsq;
    a.simpar = dev_simpar;
    a.replay = nullptr;
    const bool readout = dev_identity != nullptr || dev_rsq != nullptr;
    return launch_simulate(ctx, a, ti, readout, (int64_t)n_active * sim_count, n_active, dev_scratch, dev_scratch_bytes,
                           stream);
}

int sonics_debug_counters(sonics_ctx *ctx, uint64_t *dev_counters)
{
    if (!ctx)
        return fail(SONICS_ERR_ARG, "sonics_debug_counters: ctx is NULL");
    ctx->dbg = (unsigned long long *)dev_counters;
    return SONICS_OK;
}

int sonics_debug_trace(sonics_ctx *ctx, SonicsTraceEntry *dev_trace, int32_t trace_cap, int32_t *dev_trace_len)
{
    if (!ctx)
        return fail(SONICS_ERR_ARG, "sonics_debug_trace: ctx is NULL");
    if (dev_trace && (trace_cap <= 0 || !dev_trace_len))
        return fail(SONICS_ERR_ARG, "sonics_debug_trace: trace_cap and dev_trace_len are required with a buffer");
    ctx->trace = dev_trace;
    ctx->trace_len = dev_trace ? dev_trace_len : nullptr;
    ctx->trace_cap = dev_trace ? trace_cap : 0;
    return SONICS_OK;
}

int sonics_debug_philox(sonics_ctx *ctx, uint64_t seed, uint64_t sim, uint32_t uid, uint32_t stream_id, int n_blocks,
                        uint32_t *dev_out)
{
    if (!ctx || !dev_out || n_blocks <= 0)
        return fail(SONICS_ERR_ARG, "sonics_debug_philox: bad argument");
    CUDA_OK(cudaSetDevice(ctx->device));
    k_debug_philox<<<1, 1>>>(seed, sim, uid, stream_id, n_blocks, dev_out);
    ctx->launches++;
    CUDA_OK(cudaGetLastError());
    CUDA_OK(cudaDeviceSynchronize());
    return SONICS_OK;
}

int sonics_debug_pairs(sonics_ctx *ctx, uint64_t seed, uint64_t sim, uint32_t uid, int n_pairs, uint32_t *dev_out)
{
    if (!ctx || !dev_out || n_pairs <= 0)
        return fail(SONICS_ERR_ARG, "sonics_debug_pairs: bad argument");
    CUDA_OK(cudaSetDevice(ctx->device));
    k_debug_pairs<<<1, 1>>>(seed, sim, uid, n_pairs, dev_out);
    ctx->launches++;
    CUDA_OK(cudaGetLastError());
    CUDA_OK(cudaDeviceSynchronize());
    return SONICS_OK;
}

int sonics_debug_sample(sonics_ctx *ctx, int kind, int n, double p, int64_t count, uint64_t seed, int32_t *dev_out)
{
    if (!ctx || !dev_out || count <= 0 || kind < 0 || kind > 2)
        return fail(SONICS_ERR_ARG, "sonics_debug_sample: bad argument");
    CUDA_OK(cudaSetDevice(ctx->device));
    k_debug_sample<<<(unsigned)((count + 255) / 256), 256>>>(kind, n, p, count, seed, dev_out);
    ctx->launches++;
    CUDA_OK(cudaGetLastError());
    CUDA_OK(cudaDeviceSynchronize());
    return SONICS_OK;
}

} // extern "C"

#include "abi_stats.inl"
#include "host_io.inl"
