// Native host-side I/O of the C-ABI library (included by sonics_b200.cu; no device code here).
//
// With the Monte-Carlo loop on the GPU (~60 k genotypes/s) the reference's pure-Python ingest and row
// writer become the wall-clock bottleneck of VCF mode (~120 us per genotype for
// parse + pre-filters + formatting).  This file restates them natively:
//   sonics_vcf_parse   parse_vcf()              sonics:246-343   (REF / MOTIF from INFO, ALLREADS,
//                                                                 bp offset -> repeat count, --strict)
//                    + get_alleles()            sonics.pyx:17-31 (alleles <= 0 ignored, last entry wins)
//                    + the pre-filters of process_one_genotype()  sonics:353-419
//                                               (0-/1-allele shortcuts, auto-noise rule)
//                      straight into the CSR arrays sonics_genotype_batch consumes
//   sonics_rows_write  the header (sonics:174-207), the row format of monte_carlo()
//                      (sonics.pyx:122-134, 208-269) and "name\tblock\tref\trow\n" (sonics:437),
//                      numbers printed like Python's str(float) (shortest repr).
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <charconv>
#include <cstdlib>
#include <fstream>
#include <random>

struct sonics_vcf {
    // one entry per parsed (sample, line) genotype, in file order
    std::vector<int32_t> row_kind;   // 0 skip, 1 no_simulations row, 2 simulate
    std::vector<int32_t> row_sample; // index into samples
    std::vector<int32_t> row_block;  // index into blocks
    std::vector<int32_t> row_ref;
    std::vector<int32_t> row_one_allele; // kind 1: the single allele
    std::vector<std::string> samples, blocks;
    // CSR of the genotypes to simulate (kind 2), in file order
    std::vector<int32_t> allele_off{0}, allele, reads;
    std::vector<float> noise_coef;
    std::vector<std::string> warnings;
};

static bool last_match(const std::string &info, const char *key, const char *allowed, std::string &out)
{
    // the reference's regex ".*KEY=([allowed]*);.*": greedy prefix => the LAST occurrence of KEY whose run of
    // allowed characters is followed by ';'
    const size_t klen = strlen(key);
    size_t from = std::string::npos;
    while (true) {
        const size_t pos = info.rfind(key, from);
        if (pos == std::string::npos)
            return false;
        size_t e = pos + klen;
        while (e < info.size() && strchr(allowed, info[e]) != nullptr)
            ++e;
        if (e < info.size() && info[e] == ';') {
            out = info.substr(pos + klen, e - pos - klen);
            return true;
        }
        if (pos == 0)
            return false;
        from = pos - 1;
    }
}

static std::vector<std::string> split(const std::string &s, char sep)
{
    std::vector<std::string> out;
    size_t b = 0;
    while (true) {
        const size_t e = s.find(sep, b);
        if (e == std::string::npos) {
            out.push_back(s.substr(b));
            return out;
        }
        out.push_back(s.substr(b, e - b));
        b = e + 1;
    }
}

static bool parse_int(const std::string &s, long long &v)
{
    // Python int(): optional surrounding whitespace, optional sign, digits
    size_t b = 0, e = s.size();
    while (b < e && isspace((unsigned char)s[b]))
        ++b;
    while (e > b && isspace((unsigned char)s[e - 1]))
        --e;
    if (b == e)
        return false;
    size_t i = b;
    if (s[i] == '+' || s[i] == '-')
        ++i;
    if (i == e)
        return false;
    for (size_t j = i; j < e; ++j)
        if (!isdigit((unsigned char)s[j]))
            return false;
    v = strtoll(s.substr(b, e - b).c_str(), nullptr, 10);
    return true;
}

// Python's str(float): shortest repr, fixed notation for 1e-4 <= |v| < 1e16, else d.ddde[+-]XX
static void py_float(std::string &out, double v)
{
    if (v != v) {
        out += "nan";
        return;
    }
    if (v == HUGE_VAL || v == -HUGE_VAL) {
        out += v < 0 ? "-inf" : "inf";
        return;
    }
    char buf[64];
    auto r = std::to_chars(buf, buf + sizeof(buf), v, std::chars_format::scientific);
    std::string sci(buf, r.ptr);
    const size_t epos = sci.find('e');
    std::string mant = sci.substr(0, epos);
    const int ex = atoi(sci.c_str() + epos + 1);
    bool neg = false;
    if (!mant.empty() && mant[0] == '-') {
        neg = true;
        mant.erase(0, 1);
    }
    std::string digits;
    for (char c : mant)
        if (c != '.')
            digits += c;
    if (neg)
        out += '-';
    if (ex >= -4 && ex < 16) {
        if (ex >= 0) {
            if ((int)digits.size() <= ex + 1) {
                out += digits;
                out.append(ex + 1 - digits.size(), '0');
                out += ".0";
            } else {
                out += digits.substr(0, ex + 1);
                out += '.';
                out += digits.substr(ex + 1);
            }
        } else {
            out += "0.";
            out.append(-ex - 1, '0');
            out += digits;
        }
    } else {
        out += digits[0];
        if (digits.size() > 1) {
            out += '.';
            out += digits.substr(1);
        }
        out += 'e';
        out += ex < 0 ? '-' : '+';
        const int ax = ex < 0 ? -ex : ex;
        if (ax < 10)
            out += '0';
        out += std::to_string(ax);
    }
}

static void filter_string(std::string &out, int bits)
{
    if (bits & SONICS_FILTER_PASS) {
        out += "PASS";
        return;
    }
    if (bits & SONICS_FILTER_NO_SUCCESS) {
        out += "no_success";
        return;
    }
    bool first = true;
    auto add = [&](const char *s) {
        if (!first)
            out += ',';
        out += s;
        first = false;
    };
    if (bits & SONICS_FILTER_MWU_TEST)
        add("MWU_test");
    if (bits & SONICS_FILTER_BEST_RATIO)
        add("best_ratio");
    if (bits & SONICS_FILTER_PERCENTILE_RATIO)
        add("percentile_ratio");
}

// one verdict -> the tab-joined row monte_carlo() returns
static void format_row(std::string &out, const SonicsVerdict &v, int monte_carlo, const std::vector<std::string> &add_perc)
{
    auto genotype = [&]() {
        out += std::to_string(v.allele_lo);
        out += '/';
        out += std::to_string(v.allele_hi);
    };
    if (monte_carlo) {
        genotype();
        const double f[6] = {v.identity, v.identity_q75, v.r_squared, v.r_squared_q75, v.lnL, v.lnL_q75};
        for (double x : f) {
            out += '\t';
            py_float(out, x);
        }
        out += '\t';
        out += std::to_string(v.run_reps);
        return;
    }
    if (v.filter & SONICS_FILTER_NO_SUCCESS) {
        out += "./.\t.\t.\t.\tno_success\t.\t.\t.\t";
        out += std::to_string(v.run_reps);
        out += "\t.";
        return;
    }
    genotype();
    out += '\t';
    py_float(out, v.identity);
    out += '\t';
    py_float(out, v.r_squared);
    out += '\t';
    py_float(out, v.lnL);
    out += '\t';
    filter_string(out, v.filter);
    out += '\t';
    if (v.p_clamped)
        out += '1'; // high_pval clamped to the int 1 (sonics.pyx:225)
    else
        py_float(out, v.high_pval);
    out += '\t';
    py_float(out, v.best_lnL_ratio);
    out += '\t';
    py_float(out, v.percentile_lnL_ratio);
    out += '\t';
    out += std::to_string(v.run_reps);
    out += '\t';
    if (add_perc.empty()) {
        out += '.';
    } else {
        for (size_t i = 0; i < add_perc.size(); ++i) {
            if (i)
                out += ';';
            out += add_perc[i];
            out += '|';
            py_float(out, v.add_ratio[i]);
        }
    }
}

extern "C" {

/* A data line belongs to the byte range [byte_begin, byte_end) that holds its first byte; '#' lines are
 * header wherever they stand and are read by every range up to its own end (the sample names come from
 * "#CHROM").  byte_end < 0: to the end of the file.  Ranges that tile the file therefore parse every data
 * line exactly once, in file order -- this is how the ranks of a multi-GPU run share the ingest.  The
 * random choice of --strict 0 is keyed by (seed, offset of the line), so it does not depend on the tiling. */
int sonics_vcf_parse_range(const char *path, int64_t byte_begin, int64_t byte_end, int strict,
                           int one_allele_threshold, double noise_threshold, int monte_carlo, uint64_t seed,
                           sonics_vcf **out)
{
    if (!path || !out)
        return fail(SONICS_ERR_ARG, "sonics_vcf_parse: NULL argument");
    if (byte_begin < 0)
        return fail(SONICS_ERR_ARG, "sonics_vcf_parse: negative byte offset");
    std::ifstream in(path, std::ios::binary);
    if (!in)
        return fail(SONICS_ERR_ARG, "sonics_vcf_parse: cannot open %s", path);
    sonics_vcf *V = new sonics_vcf();
    std::string line;
    std::vector<int64_t> hist(SONICS_MAX_ALLELE);
    std::mt19937_64 rng;
    int64_t pos = 0; // offset of the first byte of the next line
    bool skipped_to_range = byte_begin == 0;
    while (true) {
        if (byte_end >= 0 && pos >= byte_end)
            break;
        if (!std::getline(in, line))
            break;
        const int64_t line_pos = pos;
        pos += (int64_t)line.size() + 1;
        if (!line.empty() && line.back() == '\r')
            line.pop_back();
        if (!line.empty() && line[0] == '#') {
            if (line.rfind("#CHROM", 0) == 0) {
                auto f = split(line, '\t');
                V->samples.assign(f.size() > 9 ? f.begin() + 9 : f.end(), f.end());
            }
            continue;
        }
        if (!skipped_to_range) {
            // first data line of the file: the header is complete; jump to the range
            skipped_to_range = true;
            if (line_pos < byte_begin) {
                in.seekg(byte_begin - 1);
                if (!std::getline(in, line)) // the rest of the line that straddles byte_begin - 1
                    break;
                pos = byte_begin + (int64_t)line.size(); // == offset after that line's '\n'
                continue;
            }
        }
        if (line.empty())
            continue;
        bool rng_ready = false; // --strict 0: seeded per line on first use
        auto f = split(line, '\t');
        if (f.size() < 9) {
            delete V;
            return fail(SONICS_ERR_ARG, "sonics_vcf_parse: line with %zu columns", f.size());
        }
        std::string sref, smotif;
        if (!last_match(f[7], "REF=", "0123456789.", sref)) {
            delete V;
            return fail(SONICS_ERR_ARG, "There's something wrong with the vcf file, expecting REF in the INFO field in "
                                        "the 8th column.");
        }
        if (!last_match(f[7], "MOTIF=", "TCGA", smotif)) {
            delete V;
            return fail(SONICS_ERR_ARG, "There's something wrong with the vcf file, expecting MOTIF in the INFO field "
                                        "in the 8th column.");
        }
        long long ref = 0;
        if (!parse_int(sref.substr(0, sref.find('.')), ref) || smotif.empty()) {
            delete V;
            return fail(SONICS_ERR_ARG, "sonics_vcf_parse: bad REF / empty MOTIF in block %s", f[0].c_str());
        }
        const long long L = (long long)smotif.size();
        const auto fmt = split(f[8], ':');
        int loc = -1;
        for (size_t i = 0; i < fmt.size(); ++i)
            if (fmt[i] == "ALLREADS") {
                loc = (int)i;
                break;
            }
        if (loc < 0) {
            delete V;
            return fail(SONICS_ERR_ARG, "sonics_vcf_parse: 'ALLREADS' is not in FORMAT of block %s", f[0].c_str());
        }
        V->blocks.push_back(f[0]);
        const int block_ix = (int)V->blocks.size() - 1;
        for (size_t k = 0; k < V->samples.size(); ++k) {
            if (9 + k >= f.size()) {
                V->warnings.push_back("Missing ALLREADS field in the input for block: " + f[0] + ", sample: " + V->samples[k]);
                continue;
            }
            const auto sub = split(f[9 + k], ':');
            if ((size_t)loc >= sub.size()) {
                V->warnings.push_back("Missing ALLREADS field in the input for block: " + f[0] + ", sample: " + V->samples[k]);
                continue;
            }
            // alleles = offset / len(motif) + ref ; support
            std::vector<long long> off_bp, sup;
            for (const auto &ent : split(sub[loc], ';')) {
                const auto pr = split(ent, '|');
                long long a = 0, s = 0;
                if (pr.size() < 2 || !parse_int(pr[0], a) || !parse_int(pr[1], s)) {
                    delete V;
                    return fail(SONICS_ERR_ARG, "sonics_vcf_parse: malformed ALLREADS entry '%s' (block %s, sample %s)",
                                ent.c_str(), f[0].c_str(), V->samples[k].c_str());
                }
                off_bp.push_back(a);
                sup.push_back(s);
            }
            bool any_partial = false;
            for (long long a : off_bp)
                any_partial |= (a % L) != 0;
            std::fill(hist.begin(), hist.end(), 0);
            bool range_error = false;
            auto put = [&](long long allele, long long s) {
                // "%i|%i" for s > 0 and a > 0, then get_alleles(): the last entry of an allele wins
                if (s > 0 && allele > 0) {
                    if (allele >= SONICS_MAX_ALLELE)
                        range_error = true;
                    else
                        hist[allele] = s;
                }
            };
            if (any_partial && strict == 2) {
                V->warnings.push_back("Partial repetition of the motif in block: " + f[0] + " sample: " + V->samples[k] + ". Excluding.");
                continue;
            }
            if (any_partial && strict == 0) {
                V->warnings.push_back("Partial repetition of the motif in block: " + f[0] + ", sample: " + V->samples[k] +
                                      ". Adding support from partial allele to one of the closest allele on random.");
                // move the support of every partial allele to floor or floor + 1 at random, then emit in allele order
                std::map<long long, long long> whole;
                for (size_t i = 0; i < off_bp.size(); ++i)
                    if (off_bp[i] % L == 0)
                        whole[off_bp[i] / L + ref] = sup[i];
                for (size_t i = 0; i < off_bp.size(); ++i)
                    if (off_bp[i] % L != 0) {
                        long long fl = off_bp[i] / L;
                        if (off_bp[i] < 0 && off_bp[i] % L != 0)
                            --fl; // floor division
                        if (!rng_ready) {
                            rng.seed(seed ^ (0x9E3779B97F4A7C15ull * (uint64_t)(line_pos + 1)));
                            rng_ready = true;
                        }
                        const long long target = fl + ref + (long long)(rng() & 1u);
                        whole[target] += sup[i];
                    }
                for (auto &kv : whole)
                    put(kv.first, kv.second);
            } else {
                if (any_partial)
                    V->warnings.push_back("Partial repetition of the motif in block: " + f[0] + " sample: " + V->samples[k] + ". Excluding allele.");
                for (size_t i = 0; i < off_bp.size(); ++i)
                    if (off_bp[i] % L == 0)
                        put(off_bp[i] / L + ref, sup[i]);
            }
            if (range_error) {
                delete V;
                return fail(SONICS_ERR_RANGE, "block %s sample %s: allele outside (0, %d)", f[0].c_str(), V->samples[k].c_str(),
                            SONICS_MAX_ALLELE);
            }
            // pre-filters of process_one_genotype (sonics:353-419)
            int n_alleles = 0, one_allele = 0;
            long long total = 0, min_reads = 0;
            for (int a = 0; a < SONICS_MAX_ALLELE; ++a)
                if (hist[a] != 0) {
                    if (n_alleles == 0 || hist[a] < min_reads)
                        min_reads = hist[a];
                    if (n_alleles == 0)
                        one_allele = a;
                    ++n_alleles;
                    total += hist[a];
                }
            int kind = 2;
            if (n_alleles == 0) {
                V->warnings.push_back("No alleles provided in input, skipping this genotype: " + f[0] + " for sample: " + V->samples[k] + ".");
                kind = 0;
            } else if (n_alleles == 1) {
                // above the threshold the reference writes a no_simulations row -- in default mode only
                kind = (total > one_allele_threshold && !monte_carlo) ? 1 : 0;
            }
            V->row_kind.push_back(kind);
            V->row_sample.push_back((int)k);
            V->row_block.push_back(block_ix);
            V->row_ref.push_back((int)ref);
            V->row_one_allele.push_back(kind == 1 ? one_allele : 0);
            if (kind == 2) {
                for (int a = 0; a < SONICS_MAX_ALLELE; ++a)
                    if (hist[a] != 0) {
                        V->allele.push_back(a);
                        V->reads.push_back((int32_t)hist[a]);
                    }
                V->allele_off.push_back((int32_t)V->allele.size());
                // auto-noise rule (sonics:407-416), the reference's fp64 expressions
                const double frac = (double)min_reads / (double)total * (double)n_alleles;
                const double noise_frac = frac / (frac + 1.0);
                const double nc = noise_frac < noise_threshold ? (double)min_reads / (double)total : 0.0;
                V->noise_coef.push_back((float)nc);
            }
        }
    }
    *out = V;
    return SONICS_OK;
}

int sonics_vcf_parse(const char *path, int strict, int one_allele_threshold, double noise_threshold, int monte_carlo,
                     uint64_t seed, sonics_vcf **out)
{
    return sonics_vcf_parse_range(path, 0, -1, strict, one_allele_threshold, noise_threshold, monte_carlo, seed, out);
}

void sonics_vcf_free(sonics_vcf *v) { delete v; }

/* what: 0 parsed genotypes (rows incl. skipped), 1 genotypes to simulate, 2 CSR entries, 3 samples, 4 blocks,
 *       5 warnings */
int64_t sonics_vcf_count(const sonics_vcf *v, int what)
{
    if (!v)
        return 0;
    switch (what) {
    case 0: return (int64_t)v->row_kind.size();
    case 1: return (int64_t)v->noise_coef.size();
    case 2: return (int64_t)v->allele.size();
    case 3: return (int64_t)v->samples.size();
    case 4: return (int64_t)v->blocks.size();
    case 5: return (int64_t)v->warnings.size();
    }
    return 0;
}

/* RNG keys of the genotypes to simulate, in CSR order: (line_base + data-line ordinal within this range) *
 * samples + sample index -- the genotype's cell in the (data line x sample) grid of the whole file, so the key
 * does not depend on how the file is cut into ranges. */
int sonics_vcf_uids(const sonics_vcf *v, int64_t line_base, uint32_t *uid_out)
{
    if (!v || !uid_out || line_base < 0)
        return fail(SONICS_ERR_ARG, "sonics_vcf_uids: bad argument");
    const int64_t ns = (int64_t)v->samples.size();
    size_t j = 0;
    for (size_t r = 0; r < v->row_kind.size(); ++r)
        if (v->row_kind[r] == 2) {
            const int64_t cell = (line_base + v->row_block[r]) * ns + v->row_sample[r];
            if (cell > 0xFFFFFFFFll)
                return fail(SONICS_ERR_RANGE, "more than 2^32 (line, sample) cells in the VCF");
            uid_out[j++] = (uint32_t)cell;
        }
    return SONICS_OK;
}

/* Number of data lines (not '#', not empty) whose first byte lies before byte_end (< 0: whole file): the
 * line_base of a range that starts at byte_end. */
int sonics_vcf_lines_before(const char *path, int64_t byte_end, int64_t *n_lines)
{
    if (!path || !n_lines)
        return fail(SONICS_ERR_ARG, "sonics_vcf_lines_before: NULL argument");
    FILE *fp = fopen(path, "rb");
    if (!fp)
        return fail(SONICS_ERR_ARG, "sonics_vcf_lines_before: cannot open %s", path);
    std::vector<char> buf(4 << 20);
    int64_t pos = 0, count = 0;
    bool at_start = true, pending_cr = false; // pending_cr: the line so far is a lone '\r'
    size_t n;
    while ((byte_end < 0 || pos < byte_end) && (n = fread(buf.data(), 1, buf.size(), fp)) > 0) {
        for (size_t i = 0; i < n && (byte_end < 0 || pos < byte_end); ++i, ++pos) {
            const char c = buf[i];
            if (at_start) {
                at_start = false;
                if (c == '\n') {
                    at_start = true;
                } else if (c == '\r') {
                    pending_cr = true;
                } else if (c != '#') {
                    ++count;
                }
            } else if (pending_cr) {
                pending_cr = false;
                if (c == '\n')
                    at_start = true; // "\r\n": an empty line
                else
                    ++count; // a line that begins with '\r' and goes on (malformed; the parser will report it)
            } else if (c == '\n') {
                at_start = true;
            }
        }
    }
    fclose(fp);
    *n_lines = count;
    return SONICS_OK;
}

/* Number of data lines whose FIRST byte lies in [byte_begin, byte_end) (byte_end < 0: end of file) -- the lines
 * sonics_vcf_parse_range(path, byte_begin, byte_end) turns into blocks.  memchr over a mapping of the file, so the
 * ranks of a multi-GPU run can count their own pieces in parallel and exchange the counts (the line_base of piece
 * k is the sum over the pieces before it) instead of each scanning the file from the start. */
int sonics_vcf_count_lines(const char *path, int64_t byte_begin, int64_t byte_end, int64_t *n_lines)
{
    if (!path || !n_lines || byte_begin < 0)
        return fail(SONICS_ERR_ARG, "sonics_vcf_count_lines: bad argument");
    *n_lines = 0;
    const int fd = open(path, O_RDONLY);
    if (fd < 0)
        return fail(SONICS_ERR_ARG, "sonics_vcf_count_lines: cannot open %s", path);
    struct stat sb;
    if (fstat(fd, &sb) != 0) {
        close(fd);
        return fail(SONICS_ERR_ARG, "sonics_vcf_count_lines: cannot stat %s", path);
    }
    const int64_t size = (int64_t)sb.st_size;
    const int64_t end = (byte_end < 0 || byte_end > size) ? size : byte_end;
    if (size == 0 || byte_begin >= end) {
        close(fd);
        return SONICS_OK;
    }
    const char *m = (const char *)mmap(nullptr, (size_t)size, PROT_READ, MAP_PRIVATE, fd, 0);
    close(fd);
    if (m == MAP_FAILED)
        return fail(SONICS_ERR_ARG, "sonics_vcf_count_lines: cannot map %s", path);
    // first line start at or after byte_begin
    int64_t p = byte_begin;
    if (p > 0 && m[p - 1] != '\n') {
        const char *nl = (const char *)memchr(m + p, '\n', (size_t)(size - p));
        p = nl ? (int64_t)(nl - m) + 1 : size;
    }
    int64_t count = 0;
    while (p < end) {
        const char c0 = m[p];
        // data line: not empty ("\n", "\r\n", a lone "\r" at the end of the file), not a '#' line
        const bool empty = c0 == '\n' || (c0 == '\r' && (p + 1 >= size || m[p + 1] == '\n'));
        if (!empty && c0 != '#')
            ++count;
        const char *nl = (const char *)memchr(m + p, '\n', (size_t)(size - p));
        if (!nl)
            break;
        p = (int64_t)(nl - m) + 1;
    }
    munmap((void *)m, (size_t)size);
    *n_lines = count;
    return SONICS_OK;
}

/* Copies the CSR of the genotypes to simulate (sizes from sonics_vcf_count) and the per-row kinds. */
int sonics_vcf_fill(const sonics_vcf *v, int32_t *allele_off, int32_t *allele, int32_t *reads, float *noise_coef,
                    int32_t *row_kind)
{
    if (!v)
        return fail(SONICS_ERR_ARG, "sonics_vcf_fill: NULL handle");
    if (allele_off)
        memcpy(allele_off, v->allele_off.data(), v->allele_off.size() * 4);
    if (allele)
        memcpy(allele, v->allele.data(), v->allele.size() * 4);
    if (reads)
        memcpy(reads, v->reads.data(), v->reads.size() * 4);
    if (noise_coef)
        memcpy(noise_coef, v->noise_coef.data(), v->noise_coef.size() * 4);
    if (row_kind)
        memcpy(row_kind, v->row_kind.data(), v->row_kind.size() * 4);
    return SONICS_OK;
}

const char *sonics_vcf_warning(const sonics_vcf *v, int64_t i)
{
    return (v && i >= 0 && i < (int64_t)v->warnings.size()) ? v->warnings[i].c_str() : "";
}

/* Writes <path>: header + one line per parsed genotype that is not skipped, in file order; verdicts[i] belongs
 * to the i-th genotype to simulate.  vcf == NULL: string mode, one row with (name, block, ".") and verdicts[0]. */
int sonics_rows_write_part(const char *path, const sonics_vcf *vcf, const SonicsVerdict *verdicts,
                           int64_t n_verdicts, int monte_carlo, const char *add_ratios, const char *name,
                           const char *block, int with_header)
{
    if (!path || (!verdicts && n_verdicts > 0))
        return fail(SONICS_ERR_ARG, "sonics_rows_write: NULL argument");
    std::vector<std::string> add_perc;
    if (add_ratios && *add_ratios)
        add_perc = split(add_ratios, ';');
    std::string out;
    out.reserve(1 << 20);
    if (with_header & 1)
        out += monte_carlo ? "sample\tblock\tref\tgenotype\tidentity\tidentity_75th_percentile\tr_squared\t"
                         "r_squared_75th_percentile\tlnL\tlnL_75th_percentile\trepeats\n"
                       : "sample\tblock\tref\tgenotype\tidentity\tr_squared\tlnL\tfilter\tMWUtes_pval\tbest_lnL_ratio\t"
                         "quartile_lnL_ratio\trepeats\tadditional_ratios\n";
    FILE *fp = fopen(path, (with_header & 2) ? "a" : "w");
    if (!fp)
        return fail(SONICS_ERR_ARG, "sonics_rows_write: cannot open %s", path);
    if (!vcf) {
        if (n_verdicts > 0) {
            out += name ? name : "sample";
            out += '\t';
            out += block ? block : "Block";
            out += "\t.\t";
            format_row(out, verdicts[0], monte_carlo, add_perc);
            out += '\n';
        }
    } else {
        int64_t vi = 0;
        for (size_t r = 0; r < vcf->row_kind.size(); ++r) {
            const int kind = vcf->row_kind[r];
            if (kind == 0)
                continue;
            out += vcf->samples[vcf->row_sample[r]];
            out += '\t';
            out += vcf->blocks[vcf->row_block[r]];
            out += '\t';
            out += std::to_string(vcf->row_ref[r]);
            out += '\t';
            if (kind == 1) {
                const std::string a = std::to_string(vcf->row_one_allele[r]);
                out += a + "/" + a + "\t.\t.\t.\tno_simulations\t.\t.\t.\t0\t.";
            } else {
                if (vi >= n_verdicts) {
                    fclose(fp);
                    return fail(SONICS_ERR_SPACE, "sonics_rows_write: %lld verdicts for more genotypes", (long long)n_verdicts);
                }
                format_row(out, verdicts[vi++], monte_carlo, add_perc);
            }
            out += '\n';
            if (out.size() > (1 << 20)) {
                fwrite(out.data(), 1, out.size(), fp);
                out.clear();
            }
        }
    }
    fwrite(out.data(), 1, out.size(), fp);
    fclose(fp);
    return SONICS_OK;
}

int sonics_rows_write(const char *path, const sonics_vcf *vcf, const SonicsVerdict *verdicts, int64_t n_verdicts,
                      int monte_carlo, const char *add_ratios, const char *name, const char *block)
{
    return sonics_rows_write_part(path, vcf, verdicts, n_verdicts, monte_carlo, add_ratios, name, block, 1);
}

/* Concatenates the part files (in the order given) into <path> and removes them. */
int sonics_rows_concat(const char *path, const char *const *parts, int n_parts)
{
    if (!path || (!parts && n_parts > 0))
        return fail(SONICS_ERR_ARG, "sonics_rows_concat: NULL argument");
    FILE *fo = fopen(path, "wb");
    if (!fo)
        return fail(SONICS_ERR_ARG, "sonics_rows_concat: cannot open %s", path);
    std::vector<char> buf(4 << 20);
    for (int i = 0; i < n_parts; ++i) {
        FILE *fi = fopen(parts[i], "rb");
        if (!fi) {
            fclose(fo);
            return fail(SONICS_ERR_ARG, "sonics_rows_concat: cannot open %s", parts[i]);
        }
        size_t n;
        while ((n = fread(buf.data(), 1, buf.size(), fi)) > 0)
            if (fwrite(buf.data(), 1, n, fo) != n) {
                fclose(fi);
                fclose(fo);
                return fail(SONICS_ERR_SPACE, "sonics_rows_concat: short write to %s", path);
            }
        fclose(fi);
        remove(parts[i]);
    }
    fclose(fo);
    return SONICS_OK;
}

} // extern "C"
