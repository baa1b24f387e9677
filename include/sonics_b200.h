/*
 * sonics_b200.h -- C-ABI of the B200-native SONiCS Monte-Carlo engine.
 *
 * This is the drop-in boundary for the reference's hot path.  In the reference
 * the CLI script imports the compiled extension module `sonics` and calls
 *     sonics.get_alleles(genotype_str)                         (sonics:350)
 *     sonics.monte_carlo(reps, constants, ranges, options)     (sonics:427-432)
 * which in turn run one_repeat()/simulate() (sonics.pyx:293-480) and the f2py
 * routine pymc_extracted.mvhyperg (pymc_extracted.f:63-109).  A maintainer
 * binds the functions below with ctypes; the Python shim
 * sonics_b200/sonics.py re-exposes get_alleles()/monte_carlo()/one_repeat()
 * with the reference's signatures on top of them.
 *
 * Conventions: extern "C", plain pointers and sizes, int status (0 = ok,
 * negative = error; sonics_last_error() gives the message of the calling
 * thread's last failure), no exceptions cross the boundary.  The library never
 * allocates device memory: every device buffer (tables, lnL arrays, workspace)
 * is owned by the caller (PyTorch tensors in the Python host) and passed as a
 * raw device pointer.  All work is enqueued on the CUDA stream handle passed in
 * (`stream`, a cudaStream_t cast to void*; NULL = legacy default stream) and is
 * asynchronous unless stated otherwise.  One ctx per GPU; a ctx is not
 * thread-safe (one host thread per GPU).
 *
 * There is no CPU fallback: every entry point that computes fails with
 * SONICS_ERR_CUDA when no sm_100 device is usable.
 */
#ifndef SONICS_B200_H
#define SONICS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SONICS_ABI_VERSION 1
#define SONICS_MAX_ALLELE 500      /* get_alleles(): max_allele, sonics.pyx:21 */
#define SONICS_SENTINEL (-999999.0) /* lnL when reads > pool, sonics.pyx:360-361 */
#define SONICS_MAX_ADD_RATIOS 8     /* --add_ratios percentiles carried per verdict */
#define SONICS_STATS_MAX_REPS 8192  /* simulations per genotype the on-chip statistics path holds */

enum {
    SONICS_OK = 0,
    SONICS_ERR_ARG = -1,     /* bad argument (null pointer, size, range) */
    SONICS_ERR_CUDA = -2,    /* CUDA runtime / no usable device */
    SONICS_ERR_ALLELES = -3, /* "Less then two alleles as starting conditions!" sonics.pyx:316-318 */
    SONICS_ERR_RANGE = -4,   /* allele index or molecule count outside the supported range */
    SONICS_ERR_SPACE = -5    /* caller-provided buffer too small */
};

/* FILTER bits of a verdict (sonics.pyx:244-252; label strings in the same order) */
enum {
    SONICS_FILTER_PASS = 1,
    SONICS_FILTER_MWU_TEST = 2,
    SONICS_FILTER_BEST_RATIO = 4,
    SONICS_FILTER_PERCENTILE_RATIO = 8,
    SONICS_FILTER_NO_SUCCESS = 16
};

typedef struct sonics_ctx sonics_ctx;

/* The `constants` / `ranges` / `options` dicts frozen into a POD
 * (sonics:714-751; keys consumed at sonics.pyx:67-82,302-337,405-410). */
typedef struct SonicsParams {
    int32_t n_cycles;      /* constants['n_cycles'] = pcr_cycles + after_capture        */
    int32_t capture_cycle; /* constants['capture_cycle'] = pcr_cycles + 1                */
    int32_t floor_opt;     /* constants['floor']; -1 => minimum repeat count 1           */
    int32_t random_start;  /* constants['random']                                        */
    int32_t up_preference; /* constants['up_preference']                                 */
    int32_t start_copies;  /* constants['start_copies']                                  */
    int32_t min_sim;       /* options['min_sim']                                         */
    int32_t monte_carlo;   /* options['monte_carlo']                                     */
    double down[2];        /* ranges[0] = (min, max)                                     */
    double up[2];          /* ranges[1]                                                  */
    double capture[2];     /* ranges[2]                                                  */
    double efficiency[2];  /* ranges[3]                                                  */
    double padjust;        /* constants['padjust']                                       */
    double lnL_threshold;  /* constants['lnL_threshold']                                 */
    int32_t n_add_ratios;  /* options['add_ratios'] split on ';' (sonics.pyx:232-242)    */
    int32_t _pad;
    double add_ratios[SONICS_MAX_ADD_RATIOS];
} SonicsParams;

/* One verdict per genotype: everything monte_carlo() prints (sonics.pyx:254-265). */
typedef struct SonicsVerdict {
    int32_t allele_lo, allele_hi; /* called genotype "lo/hi"; -1/-1 = "./."             */
    int32_t filter;               /* SONICS_FILTER_* bits                               */
    int32_t run_reps;             /* simulations run                                    */
    int32_t min_sim;              /* min over groups of #(lnL > sentinel)               */
    int32_t best_sim;             /* simulation index of the reported (best) row        */
    int32_t n_groups;
    int32_t p_clamped;            /* 1 when high_pval was clamped to the int 1 (:225)   */
    double lnL;                   /* best row                                           */
    double identity;              /* best row (replayed)                                */
    double r_squared;             /* best row (replayed), float32 widened               */
    double high_pval;             /* Bonferroni-scaled max MWU p                        */
    double best_lnL_ratio;
    double percentile_lnL_ratio;
    /* --monte_carlo extras (sonics.pyx:119-131): 75th percentiles within the best group */
    double identity_q75, r_squared_q75, lnL_q75;
    /* --add_ratios: q(best) - q(second) per requested percentile.  As in the reference the LAST of them
     * also overwrites percentile_lnL_ratio before FILTER and the output column are produced (:236-238). */
    double add_ratio[SONICS_MAX_ADD_RATIOS];
} SonicsVerdict;

int sonics_abi_version(void);
const char *sonics_last_error(void);

/* Context bound to one CUDA device (must be sm_100). */
int sonics_create(int device, sonics_ctx **out);
void sonics_destroy(sonics_ctx *ctx);
/* Block until everything enqueued on `stream` has finished. */
int sonics_sync(sonics_ctx *ctx, void *stream);

/* pymc_extracted.mvhyperg (pymc_extracted.f:63-109; f2py signature :76-79),
 * evaluated on the device in fp64 with the reference's NR-Lanczos gammln in the
 * reference's operation order.  x, color: HOST arrays of n_pools rows of k
 * int64; like_out: HOST array of n_pools doubles.  Synchronous.
 * dev_scratch: caller-owned device buffer of >= n_pools*(2*k*8 + 8) bytes. */
int sonics_mvhyperg(sonics_ctx *ctx, const int64_t *x, const int64_t *color, int k, int n_pools, double *like_out,
                    void *dev_scratch, size_t dev_scratch_bytes);

/* ---- genotype table -------------------------------------------------- */
/* A genotype (one locus x sample read-support string, get_alleles() output,
 * sonics.pyx:17-31) is given in CSR form: its non-zero alleles
 * allele[allele_off[g] .. allele_off[g+1]) ascending, with reads[] > 0.
 * noise_coef[g] is constants['noise_coef'] (sonics:407-419), applied as the
 * reference does through a C float.  uid[g] (nullable => g) keys the RNG stream
 * so results do not depend on how genotypes are sharded across GPUs. */
size_t sonics_table_bytes(int n_gt, int n_alleles_total);
int sonics_table_build(sonics_ctx *ctx, const SonicsParams *params, int n_gt, const int32_t *allele_off,
                       const int32_t *allele, const int32_t *reads, const float *noise_coef, const uint32_t *uid,
                       void *dev_table, size_t dev_table_bytes, void *stream);
/* Largest live window (histogram bins) over the table: sizes shared memory. */
int sonics_table_max_window(sonics_ctx *ctx);
/* Forget a table built into dev_table (the ctx keeps a small host-side record per table; call this before the
 * buffer is freed or reused for something else).  Unknown pointers are ignored. */
int sonics_table_release(sonics_ctx *ctx, const void *dev_table);

/* ---- K1: fixed-count simulations (one_repeat() x sim_count per genotype) --
 * For every genotype g listed in dev_active (all n_gt when dev_active is NULL)
 * run simulations j = sim_begin .. sim_begin+sim_count-1 (sonics.pyx:293-394
 * incl. simulate() :397-480 and the lnL :360-363) and store, at
 * o = g*out_stride + out_offset + (j - sim_begin)  (the repeat-until-pass loop uses out_offset = sim_begin,
 * i.e. a cumulative per-genotype array; a rank that owns a slice of the index range uses 0),
 *   dev_lnL  [o]   fp64 lnL or SONICS_SENTINEL
 *   dev_group[o]   u16 group id = lo_idx*n_alleles + hi_idx, the indices (into the
 *                                 genotype's allele list) of the two start alleles, lo_idx <= hi_idx
 * and, when the pointers are non-NULL,
 *   dev_identity (fp64), dev_rsq (fp32; a C float in the reference, sonics.pyx:300): descriptors of
 *                                 the simulated sequencing readout, sonics.pyx:341-357,365-384
 *   dev_simpar [4*o .. 4*o+3]   fp64 down, up, capture, efficiency (the report columns :387-393).
 * The RNG is counter-based (Philox4x32-10 for priors / start alleles / readout, Philox2x32-10 for the draws
 * of the PCR cycles), keyed by (seed; uid, j): results do not depend on launch geometry, on
 * sim_begin/sim_count splits, or on the GPU that runs them. */
int sonics_simulate(sonics_ctx *ctx, const SonicsParams *params, const void *dev_table, int n_gt,
                    const int32_t *dev_active, int n_active, int64_t sim_begin, int64_t sim_count, uint64_t seed,
                    int64_t out_stride, int64_t out_offset, double *dev_lnL, uint16_t *dev_group,
                    double *dev_identity, float *dev_rsq, double *dev_simpar, void *dev_scratch,
                    size_t dev_scratch_bytes, void *stream);
/* K1 exists in two forms that run the same simulations -- same draws in the same order from the same Philox
 * streams, bit-identical results (tests/test_gpu_lockstep.py):
 *   SONICS_KERNEL_LOCKSTEP    sort the work units by (genotype, start alleles, quantised priors), then 32 similar
 *                             simulations per warp in lockstep, lnL / readout fused (sim_lockstep.cuh);
 *   SONICS_KERNEL_PERSISTENT  one lane = one simulation at a time as a state machine + a separate lnL kernel
 *                             (sim_kernel.cuh); the draw trace and the control-flow counters live here.
 * SONICS_KERNEL_AUTO (default) = lockstep unless a trace / counter buffer is attached.  The environment variable
 * SONICS_K1=persistent|lockstep sets the initial value of a new ctx (measurement switch). */
enum { SONICS_KERNEL_AUTO = 0, SONICS_KERNEL_PERSISTENT = 1, SONICS_KERNEL_LOCKSTEP = 2 };
int sonics_set_kernel(sonics_ctx *ctx, int kernel);
/* Lockstep form, SONICS_SORT_SLOT_MAJOR: bits of the sort key spent on the quantised efficiency / down / up priors (each <= 15; eff_bits
 * < 0 = automatic: about log2(simulations per start-allele group / 32) bits, half of them on the efficiency).  The
 * sort only decides which lane runs which simulation: results never depend on it (tuning knob; env
 * SONICS_SORT_BITS=e,d,u sets the initial value). */
int sonics_set_sort_bits(sonics_ctx *ctx, int eff_bits, int down_bits, int up_bits);
/* Samplers of the PCR cycles (lockstep form).
 *   SONICS_SAMPLERS_EXACT (default)  inversion / BTRS: exact binomial variates (chi^2 against the pmf at 2e7 draws);
 *   SONICS_SAMPLERS_FAST             opt-in: binomial draws with n p q >= min_npq (>= 20; 100 is the tested setting)
 *                                    come from one normal deviate through the first Cornish-Fisher term, rounded
 *                                    (|cdf error| <= 1.2e-2 / (n p q)); smaller draws stay exact.  No rejection
 *                                    rounds, ~1.5x the reactions/s; passes the same KS and concordance gates
 *                                    (tests/test_gpu_fast_samplers.py) but is not the reference's sampler: results
 *                                    of the two modes differ draw by draw.  A run (and its replay, K4) must use one
 *                                    mode throughout.  The persistent / traced form is always exact. */
enum { SONICS_SAMPLERS_EXACT = 0, SONICS_SAMPLERS_FAST = 1 };
int sonics_set_samplers(sonics_ctx *ctx, int mode, double min_npq);
/* Lockstep form: which simulations share a warp.
 *   SONICS_SORT_PAIR_MAJOR (default)  noise flag | the two start alleles as absolute repeat counts | reach of the noise
 *                                     molecules below / above them | the three priors, Morton-interleaved.  A reaction
 *                                     depends on its start alleles, the noise molecules and its priors; the genotype's
 *                                     reads only enter the lnL, so like simulations of ALL genotypes of a launch are
 *                                     pooled (a genotyping round runs 100-400 simulations per genotype, 12-50 per
 *                                     start pair: less than a warp);
 *   SONICS_SORT_SLOT_MAJOR            genotype | start-allele group | eff | down | up bit fields
 *                                     (sonics_set_sort_bits).
 * Results never depend on it (env SONICS_SORT_MODE=pair|slot sets the initial value). */
enum { SONICS_SORT_PAIR_MAJOR = 0, SONICS_SORT_SLOT_MAJOR = 1 };
int sonics_set_sort_mode(sonics_ctx *ctx, int mode);
/* Scratch for sonics_simulate / sonics_replay_best.  Lockstep form: sort keys and unit order, 24 bytes per work
 * unit + CUB temporary storage.  Persistent form: the final PCR pools of the work units of one chunk (K1a -> K1b
 * hand-off, 128 * max_window bytes per 32 units).  sonics_scratch_bytes(n_units, max_window) is the size that runs
 * n_units = n_active * sim_count in a single chunk of the lockstep form, sonics_scratch_bytes_kernel(.., kernel)
 * the same for a given form; anything from the size for 32 units up works (more chunks). */
size_t sonics_scratch_bytes(int64_t n_units, int max_window);
size_t sonics_scratch_bytes_kernel(int64_t n_units, int max_window, int kernel);

/* ---- K2/K3: per-genotype statistics on the cumulative lnL arrays ---------
 * Restates sonics.pyx:136-191 (+ :208-252 for the final verdict): grouping,
 * min_sim, top-2 groups by max lnL, 75th-percentile difference, one-sided
 * Mann-Whitney U of the best group against every other group with Bonferroni,
 * PASS predicate.  n_run = cumulative simulations per genotype; up to SONICS_STATS_MAX_REPS the whole
 * reduction runs on chip (one CTA per genotype), above it through CUB sorts and scans over dev_scratch.
 * final_round != 0 also fills the FILTER bits of a non-passing genotype.  dev_identity / dev_rsq
 * are only read in --monte_carlo mode (sonics.pyx:115-131) and may be NULL otherwise.
 * Verdicts are written to dev_verdict[g] (genotype index). */
int sonics_evaluate(sonics_ctx *ctx, const SonicsParams *params, const void *dev_table, int n_gt,
                    const int32_t *dev_active, int n_active, int64_t n_run, int64_t out_stride, const double *dev_lnL,
                    const uint16_t *dev_group, const double *dev_identity, const float *dev_rsq, int final_round,
                    SonicsVerdict *dev_verdict, void *dev_scratch, size_t dev_scratch_bytes, void *stream);
/* Scratch sonics_evaluate needs when n_run > SONICS_STATS_MAX_REPS or a genotype has more than 2048 start-allele
 * groups (statistics over global memory: one segmented CUB sort, then the on-chip algorithm on global arrays, both
 * modes); 0 otherwise, where dev_scratch may be NULL.  Size for one genotype = the minimum. */
size_t sonics_stats_scratch_bytes(int64_t n_run, int max_alleles, int random_start);
/* The same for n_active genotypes in one statistics launch (26 bytes per simulation; a smaller buffer, down to the
 * one-genotype size, means several launches). */
size_t sonics_stats_scratch_bytes_batch(int n_active, int64_t n_run, int max_alleles, int random_start);

/* K4: replay simulation dev_verdict[g].best_sim of every listed genotype (same Philox stream, hence
 * the same PCR pool and lnL) with the sequencing readout switched on, and fill dev_verdict[g].identity
 * and .r_squared (sonics.pyx:341-357,365-384 for the reported row only).  Genotypes whose best_sim
 * is negative (no_success) are skipped. */
int sonics_replay_best(sonics_ctx *ctx, const SonicsParams *params, const void *dev_table, int n_gt,
                       const int32_t *dev_active, int n_active, uint64_t seed, SonicsVerdict *dev_verdict,
                       void *dev_scratch, size_t dev_scratch_bytes, void *stream);

/* ---- the whole repeat-until-pass loop for a batch of genotypes ------------
 * Equivalent of mapping process_one_genotype() -> monte_carlo() over the list
 * (sonics:216-228): rounds at 100/500/1000/5000... cumulative simulations
 * (sonics.pyx:79-82,193-199), statistics on the device after each round,
 * compaction of the genotypes that passed, replay of the best simulation for
 * identity / r^2.  Inputs and verdict_out are HOST buffers; dev_work is a
 * caller-owned device buffer of >= sonics_genotype_work_bytes(...) bytes.
 * Synchronous (returns when verdict_out is filled). */
size_t sonics_genotype_work_bytes(int n_gt, int n_alleles_total, int max_reps);
/* The same for one mode: the default mode needs no identity / r^2 planes (10 instead of 22 bytes per simulation).
 * A caller with many genotypes and a large -r picks its batch size from this (driver.run_batch_verdicts). */
size_t sonics_genotype_work_bytes_mode(int n_gt, int n_alleles_total, int max_reps, int monte_carlo);
/* ... and for the actual inputs of a call (knows the largest number of start-allele groups: --random with more
 * than 45 alleles, or -r above SONICS_STATS_MAX_REPS, route the statistics through global memory). */
size_t sonics_genotype_work_bytes_for(const SonicsParams *params, int n_gt, const int32_t *allele_off, int max_reps);
int sonics_genotype_batch(sonics_ctx *ctx, const SonicsParams *params, int n_gt, const int32_t *allele_off,
                          const int32_t *allele, const int32_t *reads, const float *noise_coef, const uint32_t *uid,
                          int max_reps, uint64_t seed, SonicsVerdict *verdict_out, void *dev_work,
                          size_t dev_work_bytes, void *stream);

/* ---- test hooks (unit tests of the device RNG and samplers) ---------------
 * sonics_debug_philox:  n_blocks x 4 u32 of the Philox4x32-10 stream (seed; sim, uid, stream): priors, start
 *                       alleles, readout.
 * sonics_debug_pairs:   n_pairs x 2 u32 of the Philox2x32-10 stream the PCR cycles draw from
 *                       (key = sim_lo ^ seed_lo, counter = (pair index | sim_hi << 24, uid ^ seed_hi)).
 * sonics_debug_sample:  count draws, one per thread (thread t uses sim = t), of
 *      kind 0: Binomial(n, p)    kind 1: Poisson(lam = p)    kind 2: Binomial(n, p) from the approximate sampler
 *      of SONICS_SAMPLERS_FAST (whatever n p q is)
 * into dev_out (int32).  Both synchronous; dev_out is caller-owned device memory. */
/* sonics_debug_counters: while dev_counters (16 x u64 of caller-owned device memory, zeroed by the
 * caller) is non-NULL, sonics_simulate accumulates K1 control-flow statistics into it: [0] passes,
 * [1..7] lanes per state-machine mode summed over passes, [8] ADVANCE inner trips, [9] lanes entering
 * ADVANCE, [10] deferred-test executions, [11] lanes served by them, [12] warps. */
int sonics_debug_counters(sonics_ctx *ctx, uint64_t *dev_counters);
/* sonics_debug_trace: while dev_trace is non-NULL, sonics_simulate runs the traced instantiation of K1a and
 * records, per simulation (unit = genotype slot * sim_count + simulation, all units must fit one scratch
 * chunk), every random draw the simulation requests, in request order, with its outcome:
 *   kind 1: Binomial(n, p) -> x   (sonics.pyx:454,458,462,478; draws with n == 0 are never requested)
 *   kind 2: Poisson(p) of a bin holding n molecules -> x, before the min(x, n) of sonics.pyx:421
 * dev_trace is [units][trace_cap] entries, dev_trace_len[unit] the number requested.  The final pool of unit
 * u stays in the caller's scratch: bin c at int32 index 64 + ((u / 32) * max_window + c) * 32 + u % 32.
 * tests/test_gpu_trace.py replays the reference's loop on the recorded outcomes (scripted draws)
 * and requires bit-identical pools.  Pass NULL to switch tracing off. */
typedef struct {
    int32_t kind;
    int32_t n;
    float p;
    int32_t x;
} SonicsTraceEntry;
int sonics_debug_trace(sonics_ctx *ctx, SonicsTraceEntry *dev_trace, int32_t trace_cap, int32_t *dev_trace_len);
/* sonics_debug_readout: while dev_readout is non-NULL, every simulation that draws a sequencing readout
 * (sonics_simulate with dev_identity / dev_rsq, sonics_replay_best) also stores it, per window bin, at
 * dev_readout[unit * max_window + c] (unit = genotype slot * sim_count + simulation; bin c = repeat count
 * window origin + c).  tests/test_gpu_readout.py feeds these vectors to the oracle's identity / rsq() restatement
 * (sonics.pyx:273-291,373-384) and requires bit-identical descriptors.  Pass NULL to switch it off. */
int sonics_debug_readout(sonics_ctx *ctx, int32_t *dev_readout);
int sonics_debug_philox(sonics_ctx *ctx, uint64_t seed, uint64_t sim, uint32_t uid, uint32_t stream_id, int n_blocks,
                        uint32_t *dev_out);
int sonics_debug_pairs(sonics_ctx *ctx, uint64_t seed, uint64_t sim, uint32_t uid, int n_pairs, uint32_t *dev_out);
int sonics_debug_sample(sonics_ctx *ctx, int kind, int n, double p, int64_t count, uint64_t seed, int32_t *dev_out);

/* ---- native host I/O (no GPU involved) ----------------------------------------------------------
 * sonics_vcf_parse restates parse_vcf() (sonics:246-343), get_alleles() (sonics.pyx:17-31) and the
 * pre-filters of process_one_genotype() (sonics:353-419: 0-/1-allele shortcuts, auto-noise rule) and
 * returns the genotypes to simulate in the CSR form sonics_genotype_batch takes.  strict as --strict
 * (the reference's two defects -- strict 0 IndexError, strict 2 column slip -- are not reproduced).
 * sonics_vcf_count(what): 0 parsed genotypes, 1 genotypes to simulate, 2 CSR entries, 3 samples,
 * 4 blocks, 5 warnings.  sonics_vcf_fill copies allele_off[n_sim+1], allele[], reads[], noise_coef[n_sim]
 * and row_kind[n_parsed] (0 skipped, 1 no_simulations row, 2 simulated); any pointer may be NULL.
 * sonics_rows_write writes the output file of run_sonics(): header (sonics:174-207) + one
 * "sample\tblock\tref\t<row>" line (sonics:437) per genotype in input order, rows formatted as
 * monte_carlo() does (sonics.pyx:122-134,208-269) with Python's str(float) number format; verdicts[i]
 * belongs to the i-th simulated genotype.  vcf == NULL: string mode (one row: name, block, ".").
 *
 * Multi-GPU runs share the host work (the reference's Pool.starmap fans out whole genotypes, sonics:216-221;
 * its single parse and per-row append, sonics:246-343,436, would bound 8 GPUs): sonics_vcf_parse_range
 * parses the data lines whose first byte lies in [byte_begin, byte_end) (byte_end < 0: end of file; header
 * lines are read by every range), so ranges that tile the file parse every line exactly once, in order;
 * sonics_rows_write_part writes one range's rows (flags: 1 = header line first, 2 = append to the file);
 * sonics_rows_concat joins the parts in rank order and removes them.  sonics_vcf_uids gives every genotype
 * to simulate its RNG key = its cell in the (data line x sample) grid of the whole file, from the number of
 * data lines before the range (sonics_vcf_lines_before, a newline scan; sonics_vcf_count(4) of the ranges
 * before it gives the same number), so results do not depend on how the file is cut.  The random choice of --strict 0 is keyed by (seed, line
 * offset) and does not depend on the tiling. */
typedef struct sonics_vcf sonics_vcf;
int sonics_vcf_parse(const char *path, int strict, int one_allele_threshold, double noise_threshold, int monte_carlo,
                     uint64_t seed, sonics_vcf **out);
int sonics_vcf_parse_range(const char *path, int64_t byte_begin, int64_t byte_end, int strict,
                           int one_allele_threshold, double noise_threshold, int monte_carlo, uint64_t seed,
                           sonics_vcf **out);
void sonics_vcf_free(sonics_vcf *vcf);
int64_t sonics_vcf_count(const sonics_vcf *vcf, int what);
int sonics_vcf_fill(const sonics_vcf *vcf, int32_t *allele_off, int32_t *allele, int32_t *reads, float *noise_coef,
                    int32_t *row_kind);
int sonics_vcf_uids(const sonics_vcf *vcf, int64_t line_base, uint32_t *uid_out);
int sonics_vcf_lines_before(const char *path, int64_t byte_end, int64_t *n_lines);
/* Data lines whose first byte lies in [byte_begin, byte_end) (< 0: end of file) = the blocks
 * sonics_vcf_parse_range(path, byte_begin, byte_end) yields; a memchr scan of that range only.  The ranks of a
 * multi-GPU run count their own pieces and exchange the counts to learn every piece's line_base. */
int sonics_vcf_count_lines(const char *path, int64_t byte_begin, int64_t byte_end, int64_t *n_lines);
const char *sonics_vcf_warning(const sonics_vcf *vcf, int64_t i);
int sonics_rows_write(const char *path, const sonics_vcf *vcf, const SonicsVerdict *verdicts, int64_t n_verdicts,
                      int monte_carlo, const char *add_ratios, const char *name, const char *block);
int sonics_rows_write_part(const char *path, const sonics_vcf *vcf, const SonicsVerdict *verdicts,
                           int64_t n_verdicts, int monte_carlo, const char *add_ratios, const char *name,
                           const char *block, int flags);
int sonics_rows_concat(const char *path, const char *const *parts, int n_parts);

/* Kernel launches issued by this ctx so far (bench.py's gpu_launches). */
int64_t sonics_launch_count(sonics_ctx *ctx);

#ifdef __cplusplus
}
#endif
#endif /* SONICS_B200_H */
