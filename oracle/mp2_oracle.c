/*
 * mp2_oracle.c -- CPU restatement of MAGpurify2's composition / coverage hot paths.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under magpurify2_b200/ may import, link or
 * execute this file; only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs use it, and only as the checker / timed
 * CPU baseline.  It is a loop-for-loop restatement of the reference algorithm,
 * NOT the reference binary (the Rust crate cannot be built here: no cargo/rustc).
 *
 * Pinning status
 *   composition (get_tnf / get_gc)      : restated from /root/reference/src/composition.rs,
 *                                         LUT compared byte-for-byte with the table in that
 *                                         file at test time; KATs in tests/golden/.
 *   weighted median / log-ratio scores  : pinned by golden vectors produced by the reference's
 *                                         own NumPy functions (tests/golden/make_golden.py).
 *   coverage (coverm 0.3.2 arithmetic)  : PARITY UNPINNED.  coverm is an un-vendored Cargo
 *                                         dependency (Cargo.toml:12, no lockfile); restated from
 *                                         SURVEY.md Appendix A (recalled upstream behaviour) and
 *                                         anchored on the call site src/coverage.rs:99-142.
 *
 * Build: see oracle/Makefile  (gcc -O3 -fopenmp -ffp-contract=off -shared -fPIC).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define MP2O_NCANON 136

/* ------------------------------------------------------------------------- */
/* A1: REV_COMPLEMENTARY_FOURMERS (src/composition.rs:24-37).                 */
/* Regenerated from its rule instead of copied: code = 2 bits per base        */
/* (A=0,C=1,G=2,T=3, first base most significant); class = rank of            */
/* min(code, revcomp(code)) among the 136 canonical codes in ascending order. */
/* ------------------------------------------------------------------------- */
static unsigned revcomp4(unsigned c) {
    unsigned r = 0;
    for (int i = 0; i < 4; i++) {
        r = (r << 2) | (3u - (c & 3u));
        c >>= 2;
    }
    return r;
}

void mp2o_canonical_lut(uint8_t lut[256]) {
    int rank[256];
    int next = 0;
    for (unsigned c = 0; c < 256; c++) {
        unsigned rc = revcomp4(c);
        rank[c] = (c <= rc) ? next++ : -1;
    }
    for (unsigned c = 0; c < 256; c++) {
        unsigned rc = revcomp4(c);
        lut[c] = (uint8_t)rank[c <= rc ? c : rc];
    }
}

/* ------------------------------------------------------------------------- */
/* A2: count_kmers(seq, 4)  (src/composition.rs:42-63)                        */
/* ------------------------------------------------------------------------- */
void mp2o_count_kmers4(const uint8_t *seq, int64_t len, uint32_t counts[256]) {
    memset(counts, 0, 256 * sizeof(uint32_t));
    const uint32_t k = 4;
    uint32_t kmer = 0;
    uint32_t countdown = k - 1;                 /* :45 */
    const uint32_t mask = (1u << (2 * k)) - 1;  /* :47 */
    for (int64_t i = 0; i < len; i++) {
        switch (seq[i]) {                       /* :49-55 */
        case 65: case 97:  kmer = ((kmer << 2) | 0) & mask; break;
        case 67: case 99:  kmer = ((kmer << 2) | 1) & mask; break;
        case 71: case 103: kmer = ((kmer << 2) | 2) & mask; break;
        case 84: case 116: kmer = ((kmer << 2) | 3) & mask; break;
        default: countdown = k; break;
        }
        if (countdown == 0) counts[kmer] += 1;  /* :56-60 */
        else countdown -= 1;
    }
}

/* A3: count_canonical_fourmers  (src/composition.rs:67-74) */
void mp2o_count_canonical(const uint8_t *seq, int64_t len, uint32_t out[MP2O_NCANON]) {
    uint8_t lut[256];
    uint32_t counts[256];
    mp2o_canonical_lut(lut);
    mp2o_count_kmers4(seq, len, counts);
    memset(out, 0, MP2O_NCANON * sizeof(uint32_t));
    for (int i = 0; i < 256; i++) out[lut[i]] += counts[i];
}

/* Row sum in f32 in the order ndarray 0.13 uses for a contiguous lane
 * (numeric_util::unrolled_fold: 8 partial sums, combined (p0+p4),(p1+p5),(p2+p6),(p3+p7)).
 * Only observable when a row holds >= 2^24 k-mers; recalled, not verifiable here. */
float mp2o_rowsum_f32(const float *x, int n) {
    float p[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    int i = 0;
    for (; i + 8 <= n; i += 8)
        for (int j = 0; j < 8; j++) p[j] = p[j] + x[i + j];
    float acc = 0.0f;
    acc = acc + (p[0] + p[4]);
    acc = acc + (p[1] + p[5]);
    acc = acc + (p[2] + p[6]);
    acc = acc + (p[3] + p[7]);
    for (; i < n; i++) acc = acc + x[i];
    return acc;
}

/* A4: get_tnf  (src/composition.rs:94-124).  bases = all sequences concatenated,
 * offsets[n+1].  Parallel-for over sequences only (rayon par_iter analogue, :100-110);
 * there is no intra-sequence parallelism in the reference.  counts and/or rel may be NULL. */
int mp2o_get_tnf(const uint8_t *bases, const int64_t *offsets, int64_t n, int threads,
                 uint32_t *counts, float *rel) {
    if (threads < 1) threads = 1;
    uint8_t lut[256];
    mp2o_canonical_lut(lut);
#pragma omp parallel for schedule(dynamic, 1) num_threads(threads)
    for (int64_t s = 0; s < n; s++) {
        uint32_t c256[256], c136[MP2O_NCANON];
        mp2o_count_kmers4(bases + offsets[s], offsets[s + 1] - offsets[s], c256);
        memset(c136, 0, sizeof c136);
        for (int i = 0; i < 256; i++) c136[lut[i]] += c256[i];
        if (counts) memcpy(counts + s * MP2O_NCANON, c136, sizeof c136);
        if (rel) { /* :111-118: u32 -> f32, row sum in f32, divide, NaN -> 0 */
            float f[MP2O_NCANON];
            for (int i = 0; i < MP2O_NCANON; i++) f[i] = (float)c136[i];
            float sum = mp2o_rowsum_f32(f, MP2O_NCANON);
            for (int i = 0; i < MP2O_NCANON; i++) {
                float q = f[i] / sum;
                rel[s * MP2O_NCANON + i] = isnan(q) ? 0.0f : q;
            }
        }
    }
    return 0;
}

/* A5: gc_content / get_gc  (src/composition.rs:128-132, 148-161).  Serial, as written.
 * The reference makes five passes (4x matches + chars().count()); `passes5` != 0 keeps that
 * memory behaviour for the timed baseline, 0 uses one fused pass (same result).
 * chars().count() counts UTF-8 scalar values = bytes that are not continuation bytes. */
int mp2o_get_gc(const uint8_t *bases, const int64_t *offsets, int64_t n, int passes5, float *gc) {
    for (int64_t s = 0; s < n; s++) {
        const uint8_t *p = bases + offsets[s];
        int64_t len = offsets[s + 1] - offsets[s];
        uint64_t g = 0, c = 0, chars = 0;
        if (passes5) {
            uint64_t a = 0, b = 0, d = 0, e = 0;
            for (int64_t i = 0; i < len; i++) a += (p[i] == 'G');
            for (int64_t i = 0; i < len; i++) b += (p[i] == 'g');
            for (int64_t i = 0; i < len; i++) d += (p[i] == 'C');
            for (int64_t i = 0; i < len; i++) e += (p[i] == 'c');
            for (int64_t i = 0; i < len; i++) chars += ((p[i] & 0xC0) != 0x80);
            g = a + b; c = d + e;
        } else {
            for (int64_t i = 0; i < len; i++) {
                uint8_t ch = p[i];
                g += (ch == 'G') | (ch == 'g');
                c += (ch == 'C') | (ch == 'c');
                chars += ((ch & 0xC0) != 0x80);
            }
        }
        gc[s] = (float)(g + c) / (float)chars; /* empty sequence -> NaN, not cleaned (:131) */
    }
    return 0;
}

/* ------------------------------------------------------------------------- */
/* A7: coverm 0.3.2 trimmed-mean estimator.     */
/* PARITY UNPINNED: crate source is not under /root/reference.                 */
/* ------------------------------------------------------------------------- */
typedef struct {
    uint64_t observed_len;  /* N = len - 2*excl (0 if the contig is too short)       */
    uint64_t covered;       /* positions in the window with depth > 0                */
    uint64_t min_index;     /* floor(lo * (f32)N)                                    */
    uint64_t max_index;     /* ceil(hi * (f32)N), hi = 1.0f - trim_upper             */
    uint64_t total;         /* integer numerator of the trimmed mean                 */
    uint64_t max_depth;     /* highest occupied histogram bin                        */
    float coverage;         /* (f32)total / (f32)(max_index - min_index)             */
} mp2o_cov_stats;

/* A.4 calculate_coverage: histogram walk.  `trim_hi` is ALREADY 1.0f - trim_upper
 * (src/coverage.rs:99).  counts[d] = number of window positions with depth d. */
void mp2o_trimmed_mean_from_hist(const uint32_t *counts, uint64_t nbins, uint64_t N,
                                 uint64_t covered, float trim_lo, float trim_hi,
                                 mp2o_cov_stats *st) {
    st->observed_len = N;
    st->covered = covered;
    st->min_index = st->max_index = st->total = 0;
    st->coverage = 0.0f;
    if (N == 0) return;
    /* min_fraction_covered_bases = 0 (src/coverage.rs:100): the `< 0.0` test never fires */
    volatile float fn = (float)N;
    volatile float plo = trim_lo * fn, phi = trim_hi * fn; /* single-rounded f32 products */
    uint64_t min_index = (uint64_t)floorf(plo);
    uint64_t max_index = (uint64_t)ceilf(phi);
    st->min_index = min_index;
    st->max_index = max_index;
    if (covered == 0) return;
    uint64_t acc = 0, total = 0;
    int started = 0;
    for (uint64_t i = 0; i < nbins; i++) {
        acc += counts[i];
        if (!started) {
            if (acc >= min_index) {
                if (acc > max_index) total = (max_index - min_index + 1) * i;
                else total = (acc - min_index + 1) * i;
                started = 1;
            }
        } else {
            if (acc > max_index) {
                uint64_t excess = acc - counts[i];
                uint64_t wanted = (max_index >= excess) ? (max_index - excess + 1) : 0;
                total += wanted * i;
                break;
            } else {
                total += (uint64_t)counts[i] * i;
            }
        }
    }
    st->total = total;
    st->coverage = (float)total / (float)(max_index - min_index);
}

/* A.2 + A.3 for ONE contig of one BAM: events are the M/=/X blocks (0-based start,
 * exclusive end) of the reads that passed the filters, in any order.
 * ud[start] += 1; if end < len { ud[end] -= 1 }; prefix sum; windowed depth histogram. */
int mp2o_cov_contig(const uint32_t *starts, const uint32_t *ends, int64_t n_events,
                    uint64_t len, uint32_t excl, float trim_lo, float trim_hi,
                    mp2o_cov_stats *st) {
    memset(st, 0, sizeof *st);
    if (2ull * excl >= len) return 0; /* too short: observed length stays 0 -> 0.0 */
    int32_t *ud = (int32_t *)calloc(len, sizeof(int32_t));
    if (!ud) return -1;
    for (int64_t e = 0; e < n_events; e++) {
        if (starts[e] >= len) continue; /* never produced by a valid BAM */
        ud[starts[e]] += 1;
        if (ends[e] < len) ud[ends[e]] -= 1;
    }
    uint64_t nb = 1024;
    uint32_t *counts = (uint32_t *)calloc(nb, sizeof(uint32_t));
    uint64_t covered = 0, maxd = 0;
    int32_t cum = 0;
    uint64_t lo = excl, hi = len - excl - 1;
    for (uint64_t i = 0; i < len; i++) {
        cum += ud[i];
        if (i >= lo && i <= hi) {
            if (cum > 0) covered++;
            uint64_t d = (uint64_t)(cum < 0 ? 0 : cum);
            if (d >= nb) {
                uint64_t nn = nb;
                while (d >= nn) nn *= 2;
                counts = (uint32_t *)realloc(counts, nn * sizeof(uint32_t));
                memset(counts + nb, 0, (nn - nb) * sizeof(uint32_t));
                nb = nn;
            }
            counts[d]++;
            if (d > maxd) maxd = d;
        }
    }
    mp2o_trimmed_mean_from_hist(counts, maxd + 1, len - 2ull * excl, covered, trim_lo, trim_hi, st);
    st->max_depth = maxd;
    free(counts);
    free(ud);
    return 0;
}

/* Whole sample: events grouped by contig (ev_off[n_contigs+1]); contig_len[n_contigs].
 * Single-threaded per sample as coverm's record loop is (Appendix A.2).  `threads` > 1 is a
 * courtesy for the timed baseline only (parallel over contigs); results are identical. */
int mp2o_cov_sample(const uint32_t *starts, const uint32_t *ends, const int64_t *ev_off,
                    const int64_t *contig_len, int64_t n_contigs, uint32_t excl, float trim_lo,
                    float trim_upper, int threads, float *cov, mp2o_cov_stats *stats) {
    float trim_hi = 1.0f - trim_upper; /* src/coverage.rs:99 */
    int rc = 0;
    if (threads < 1) threads = 1;
#pragma omp parallel for schedule(dynamic, 16) num_threads(threads)
    for (int64_t c = 0; c < n_contigs; c++) {
        mp2o_cov_stats st;
        int r = mp2o_cov_contig(starts + ev_off[c], ends + ev_off[c], ev_off[c + 1] - ev_off[c],
                                (uint64_t)contig_len[c], excl, trim_lo, trim_hi, &st);
        if (r) rc = r;
        if (cov) cov[c] = st.coverage;
        if (stats) stats[c] = st;
    }
    return rc;
}

/* A.1 per-record filter + A.2 CIGAR walk for one BAM record.
 * cigar[i] = len << 4 | op  (BAM encoding: M0 I1 D2 N3 S4 H5 P6 =7 X8).
 * Returns the number of (start,end) blocks written (0 if the read is dropped),
 * or -1 when the NM tag is missing on a record that reaches the identity filter
 * (the reference panics there). */
int mp2o_record_events(uint16_t flag, int32_t tid, int32_t pos, const uint32_t *cigar,
                       int n_cigar, int64_t nm /* <0 = tag missing */, float min_identity,
                       uint32_t *out_start, uint32_t *out_end, int max_out) {
    if (flag & 0x100) return 0; /* secondary   (src/coverage.rs:104) */
    if (flag & 0x800) return 0; /* supplementary (:105)              */
    /* include_improper_pairs = true (:103): no proper-pair requirement */
    if ((flag & 0x4) || tid < 0) return 0; /* unmapped: contributes nothing */
    if (min_identity > 0.0f) {
        if (nm < 0) return -1;
        uint32_t aligned = 0;
        for (int i = 0; i < n_cigar; i++) {
            uint32_t op = cigar[i] & 0xF, l = cigar[i] >> 4;
            if (op == 0 || op == 1 || op == 7 || op == 8) aligned += l; /* M I = X */
        }
        volatile float ident = 1.0f - (float)nm / (float)aligned;
        if (!(ident >= min_identity)) return 0;
    }
    int n = 0;
    uint32_t cursor = (uint32_t)pos;
    for (int i = 0; i < n_cigar; i++) {
        uint32_t op = cigar[i] & 0xF, l = cigar[i] >> 4;
        if (op == 0 || op == 7 || op == 8) { /* M = X */
            if (n < max_out) { out_start[n] = cursor; out_end[n] = cursor + l; }
            n++;
            cursor += l;
        } else if (op == 2 || op == 3) { /* D N */
            cursor += l;
        } /* I S H P: nothing */
    }
    return n;
}

/* ------------------------------------------------------------------------- */
/* A8: tools.get_weighted_median / get_log_ratio_scores                        */
/* (magpurify2/tools.py:283-311, 314-351).                                     */
/* ------------------------------------------------------------------------- */
typedef struct { float v; int64_t w; } vw_t;
static int vw_cmp(const void *a, const void *b) {
    const vw_t *x = (const vw_t *)a, *y = (const vw_t *)b;
    if (x->v < y->v) return -1;
    if (x->v > y->v) return 1;
    if (x->w < y->w) return -1;
    if (x->w > y->w) return 1;
    return 0;
}

float mp2o_weighted_median(const float *data, const int64_t *weights, int64_t n) {
    if (n <= 0) return NAN;
    vw_t *s = (vw_t *)malloc(n * sizeof(vw_t));
    int64_t total = 0, wmax = 0;
    for (int64_t i = 0; i < n; i++) {
        s[i].v = data[i]; s[i].w = weights[i];
        total += weights[i];
        if (weights[i] > wmax) wmax = weights[i];
    }
    float med;
    if (2 * wmax > total) { /* :302-303: a single weight above the midpoint wins */
        med = NAN;
        for (int64_t i = 0; i < n; i++)
            if (weights[i] == wmax) { med = data[i]; break; }
    } else {
        qsort(s, n, sizeof(vw_t), vw_cmp); /* sorted(zip(data, weights)) :300 */
        int64_t cs = 0, idx = -1, csidx = 0;
        for (int64_t i = 0; i < n; i++) {
            cs += s[i].w;
            if (2 * cs <= total) { idx = i; csidx = cs; }
        }
        if (idx < 0) { idx = 0; csidx = -1; } /* unreachable when weights >= 0 */
        if (2 * csidx == total) {              /* :307-308: mean of the two straddlers (f32) */
            if (idx + 1 < n) med = (s[idx].v + s[idx + 1].v) / 2.0f;
            else med = s[idx].v;
        } else {
            med = s[idx + 1 < n ? idx + 1 : idx].v; /* :310 */
        }
    }
    free(s);
    return med;
}

/* data is row-major [n, s]; one weighted median per column; score = row mean of
 * 1 - |ln((x+pseudo)/(m+pseudo)) / ln(base)|, clipped to [0,1].
 * dtype flow follows NumPy >= 2: ratio and log in f32, then f64. */
int mp2o_log_ratio_scores(const float *data, const int64_t *weights, int64_t n, int64_t s,
                          double base, double *scores, float *medians_out) {
    float *col = (float *)malloc((n > 0 ? n : 1) * sizeof(float));
    float *med = (float *)malloc((s > 0 ? s : 1) * sizeof(float));
    for (int64_t j = 0; j < s; j++) {
        for (int64_t i = 0; i < n; i++) col[i] = data[i * s + j];
        med[j] = mp2o_weighted_median(col, weights, n);
        if (medians_out) medians_out[j] = med[j];
    }
    const float pseudo = 1e-5f;
    const double lb = log(base);
    for (int64_t i = 0; i < n; i++) {
        double acc = 0.0;
        for (int64_t j = 0; j < s; j++) {
            volatile float num = data[i * s + j] + pseudo;
            volatile float den = med[j] + pseudo;
            volatile float ratio = num / den;
            float lg = logf(ratio);
            acc += 1.0 - fabs((double)lg / lb);
        }
        double sc = (s == 1) ? acc : acc / (double)s;
        if (sc < 0.0) sc = 0.0;
        if (sc > 1.0) sc = 1.0;
        scores[i] = sc; /* NaN stays NaN, as np.clip does */
    }
    free(col);
    free(med);
    return 0;
}

/* core.Coverage sample selection (magpurify2/core.py:301-305):
 * np.average(cov, axis=0, weights=lengths) >= min_average_coverage, in f64. */
int mp2o_select_samples(const float *cov, const int64_t *lengths, int64_t n, int64_t s,
                        double min_average_coverage, uint8_t *selected, double *means) {
    double wsum = 0.0;
    for (int64_t i = 0; i < n; i++) wsum += (double)lengths[i];
    for (int64_t j = 0; j < s; j++) {
        double acc = 0.0;
        for (int64_t i = 0; i < n; i++) acc += (double)cov[i * s + j] * (double)lengths[i];
        double m = acc / wsum;
        if (means) means[j] = m;
        selected[j] = (m >= min_average_coverage) ? 1 : 0;
    }
    return 0;
}

/* ------------------------------------------------------------------------- */
/* A9 (self-specified, no reference code exists): CLR + centroid distance.     */
/* See DESIGN.md "tnf_score".  counts: u32[n,136] of ONE MAG.                  */
/*   clr_ij = ln(c_ij + 1) - mean_j ln(c_ij + 1)                 (f64)         */
/*   centroid_j = sum_i len_i * clr_ij / sum_i len_i                            */
/*   d_i = sqrt(sum_j (clr_ij - centroid_j)^2)                                  */
/* ------------------------------------------------------------------------- */
int mp2o_clr_centroid(const uint32_t *counts, const int64_t *lengths, int64_t n, float *clr,
                      double *dist) {
    double cen[MP2O_NCANON];
    memset(cen, 0, sizeof cen);
    double *row = (double *)malloc(sizeof(double) * MP2O_NCANON * (n > 0 ? n : 1));
    double wsum = 0.0;
    for (int64_t i = 0; i < n; i++) {
        double m = 0.0;
        for (int j = 0; j < MP2O_NCANON; j++) {
            row[i * MP2O_NCANON + j] = log((double)counts[i * MP2O_NCANON + j] + 1.0);
            m += row[i * MP2O_NCANON + j];
        }
        m /= MP2O_NCANON;
        for (int j = 0; j < MP2O_NCANON; j++) {
            row[i * MP2O_NCANON + j] -= m;
            cen[j] += (double)lengths[i] * row[i * MP2O_NCANON + j];
            if (clr) clr[i * MP2O_NCANON + j] = (float)row[i * MP2O_NCANON + j];
        }
        wsum += (double)lengths[i];
    }
    for (int j = 0; j < MP2O_NCANON; j++) cen[j] = wsum > 0 ? cen[j] / wsum : 0.0;
    for (int64_t i = 0; i < n; i++) {
        double d2 = 0.0;
        for (int j = 0; j < MP2O_NCANON; j++) {
            double d = row[i * MP2O_NCANON + j] - cen[j];
            d2 += d * d;
        }
        dist[i] = sqrt(d2);
    }
    free(row);
    return 0;
}

int mp2o_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
