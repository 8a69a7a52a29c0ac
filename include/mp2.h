/*
 * mp2.h -- C ABI of libmp2.so: B200 (sm_100a) kernels for MAGpurify2's composition and
 * coverage scoring paths.  This is the drop-in boundary: each entry point replaces the compute
 * behind one function the reference exposes through PyO3 (`magpurify2._composition`,
 * `magpurify2._coverage`, registered in /root/reference/src/lib.rs:19-22) or the NumPy scoring
 * that sits directly on their outputs.  Plain pointers and sizes only; no torch types.
 *
 * Conventions
 *   - every function returns 0 on success and a negative MP2_E* code on failure; the message is
 *     available from mp2_last_error() (thread-local).  The reference panics instead
 *     (unwrap/expect: src/composition.rs:99,109; src/coverage.rs:164,200).
 *   - "*_device" functions take DEVICE pointers (e.g. torch.Tensor.data_ptr()) and a
 *     cudaStream_t passed as void* (NULL = legacy default stream).  They are asynchronous
 *     with respect to the host: work is enqueued on `stream`, temporaries are stream-ordered.
 *   - host-pointer functions stage H2D/D2H themselves (pinned ring + async copies overlapped
 *     with the kernels) and return when the outputs are complete.
 *   - inputs are borrowed for the duration of the call; outputs are caller-allocated.
 *   - there is NO CPU fallback: without a usable CUDA device every compute call fails with
 *     MP2_ENODEV.
 */
#ifndef MP2_H
#define MP2_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MP2_NCANON 136 /* canonical 4-mer classes (src/composition.rs:67-74) */

#define MP2_OK 0
#define MP2_EINVAL -1   /* bad argument                                    */
#define MP2_ENODEV -2   /* no CUDA device / driver                         */
#define MP2_ECUDA -3    /* CUDA runtime error (message has the details)    */
#define MP2_ENOMEM -4   /* host or device allocation failed                */
#define MP2_EOVERFLOW -5 /* a per-base depth exceeded the supported range  */
#define MP2_EIO -6      /* unreadable / malformed BAM or FASTA             */
#define MP2_EFORMAT -7  /* record violates an assumption (e.g. NM missing) */

/* ---- library ------------------------------------------------------------------------ */
const char *mp2_last_error(void);
const char *mp2_version(void);
int mp2_device_count(void); /* >= 0, or a negative MP2_E* code */

/* The 256-entry code -> canonical-class table (A=0,C=1,G=2,T=3, first base most significant),
 * i.e. REV_COMPLEMENTARY_FOURMERS (src/composition.rs:24-37), regenerated from its rank rule. */
int mp2_canonical_lut(uint8_t lut[256]);

/* Per-kernel device timing for bench.py's roofline line.  When enabled, every kernel the library
 * launches is bracketed by CUDA events on its own stream (no extra synchronisation).
 * mp2_profile_read synchronises the pending events and returns the accumulated device time and
 * launch count of the kernels whose name starts with `prefix` ("" = all).  mp2_launch_count is
 * always maintained (total kernel launches by this library in this process). */
int mp2_profile_enable(int on);
int mp2_profile_reset(void);
int mp2_profile_read(const char *prefix, double *ms, int64_t *launches);
int64_t mp2_launch_count(void);

/* ---- composition: replaces _composition.get_tnf / get_gc ------------------------------ */
/* (src/composition.rs:94-124 and :148-161; call sites magpurify2/modules/composition.py:64,78) */

/* All sequences concatenated in `bases` (any bytes; only AaCcGgTt form k-mers), sequence i is
 * bases[offsets[i] .. offsets[i+1]).  Any of counts / rel / gc may be NULL.
 *   counts : uint32[n*136]  raw canonical 4-mer counts            (get_tnf(relative=False))
 *   rel    : float32[n*136] counts / row sum, NaN -> 0            (get_tnf(relative=True))
 *   gc     : float32[n]     (#G+#g+#C+#c) / byte length, 0/0=NaN  (get_gc)                */
int mp2_tnf(const uint8_t *bases, const int64_t *offsets, int64_t n, int device,
            uint32_t *counts, float *rel, float *gc);

/* Same, inputs and outputs resident on the current device.  d_counts is REQUIRED (it is the
 * accumulation target; the call defines every row of it whatever it held before: rows of contigs that
 * lie inside one piece of the stream are stored, the others are cleared first); d_rel / d_gc may be NULL.
 * n_bases = offsets[n] (passed so the call needs no device->host read). */
int mp2_tnf_device(const void *d_bases, const void *d_offsets, int64_t n, int64_t n_bases,
                   void *d_counts, void *d_rel, void *d_gc, void *stream);

/* Relative profiles from raw counts already on the device: rel[i][j] = (float)counts[i][j] / (float)row sum, NaN -> 0
 * (src/composition.rs:111-118; what get_tnf(relative=True) adds to get_tnf(relative=False)).  d_rel may alias nothing
 * of d_counts.  Lets a caller that asked mp2_tnf_device for counts only normalise beside its other work. */
int mp2_tnf_normalize_device(const void *d_counts, int64_t n, void *d_rel, void *stream);

/* 2-bit packed variant of the same path (north_star layout): codes = 2 bits per base, 16 bases
 * per uint32, FIRST base in the MOST significant bits; valid = 1 bit per base, 32 bases per
 * uint32, first base most significant (0 = N / IUPAC / any non-ACGT byte); gcmask is not needed:
 * GC is derived from the codes under the validity mask (lower-case folds to the same code).
 * n_bases positions; d_offsets as above (in bases). */
int mp2_tnf_packed_device(const void *d_codes, const void *d_valid, const void *d_offsets,
                          int64_t n, int64_t n_bases, void *d_counts, void *d_rel, void *d_gc,
                          void *stream);

/* ASCII -> (codes, valid) packer used to produce the layout above on the device. */
int mp2_pack_device(const void *d_bases, int64_t n_bases, void *d_codes, void *d_valid,
                    void *stream);

/* The same layout produced on the HOST (multi-threaded; AVX-512BW when the CPU has it), from the virtual
 * concatenation of `n_parts` byte buffers -- the genomes of a run as the FASTA reader left them (replaces the
 * per-contig Python str hand-over of magpurify2/core.py:35-50, modules/composition.py:63-66).  codes and valid
 * must hold mp2_packed_words() words (whole quads of 64 bases: 4 code words + 2 validity words each); pinned
 * memory makes the upload a plain DMA.  Returns MP2_OK / MP2_EINVAL (no message: this function has no CUDA in it). */
int64_t mp2_packed_words(int64_t n_bases, int64_t *valid_words); /* returns the number of code words */
int mp2_pack_host(const uint8_t *const *parts, const int64_t *part_len, int64_t n_parts, uint32_t *codes,
                  uint32_t *valid, int threads);

/* Runs of invalid bases of a validity mask as ascending (begin, end) pairs -- a few thousand pairs for a real
 * genome, where the mask itself is 0.125 bytes per base.  Returns the number of runs, -(needed) when `cap` pairs
 * are too few, or MP2_EINVAL. */
int64_t mp2_invalid_runs(const uint32_t *valid, int64_t n_bases, int64_t *runs, int64_t cap, int threads);

/* mp2_tnf on the packed host arrays (offsets in bases, as above; n_bases = offsets[n]).  Give EITHER the validity
 * mask (0.375 bytes per base cross PCIe in all) OR, with valid = NULL, the runs of invalid bases (0.25 bytes per
 * base: the device rebuilds the mask).  The rows of finished contigs are copied back while later slices are still
 * being uploaded.  Same outputs as mp2_tnf, any of them may be NULL. */
int mp2_tnf_packed(const uint32_t *codes, const uint32_t *valid, const int64_t *invalid_runs, int64_t n_runs,
                   const int64_t *offsets, int64_t n, int device, uint32_t *counts, float *rel, float *gc);

/* ---- coverage: replaces the coverm arithmetic behind _coverage.get_bam_coverages ------- */
/* (src/coverage.rs:99-142; coverm 0.3.2 contig_coverage / add_contig / calculate_coverage,
 *  SURVEY.md Appendix A).  One call = ONE sample (one BAM).                                */

/* Per-contig integers exposed so depth sums can be checked bit-exactly. */
typedef struct {
    uint64_t observed_len; /* len - 2*end_excl, 0 when the contig is too short */
    uint64_t covered;      /* window positions with depth > 0                   */
    uint64_t min_index;    /* floor(trim_lower * (f32)N)                        */
    uint64_t max_index;    /* ceil((1.0f - trim_upper) * (f32)N)                */
    uint64_t total;        /* integer numerator of the trimmed mean             */
    uint64_t max_depth;    /* highest occupied depth                            */
    float coverage;        /* (f32)total / (f32)(max_index - min_index)         */
    uint32_t flags;        /* MP2_COV_GENERAL_PATH when the general path computed it */
} mp2_cov_stats;

#define MP2_COV_GENERAL_PATH 1u /* mp2_cov_stats.flags: difference array went through HBM (fallback) */

/* Alignment blocks (the M/=/X runs of the reads that passed the filters), grouped by contig
 * in BAM order: block e of contig c is [starts[e], ends[e]) for ev_off[c] <= e < ev_off[c+1],
 * 0-based start, exclusive end, relative to the contig (any order inside a contig).
 * contig_off: int64[n_contigs+1] cumulative contig lengths (the same array mp2_tnf takes);
 * n_positions = contig_off[n_contigs] (passed so the call needs no device->host read).
 *   d_cov   : float32, element c at d_cov[c*cov_stride] -- cov_stride = n_bams writes column s of
 *             the reference's float32[n_contigs, n_bams] matrix in place        (REQUIRED)
 *   d_stats : mp2_cov_stats[n_contigs]    or NULL
 * max_span, max_disorder: what the caller knows about the order of the blocks inside a contig.
 *           max_span > 0 = the caller guarantees that end - start <= max_span for every block and that
 *           no start lies more than max_disorder below the largest start before it in the same contig
 *           (0 = starts non-decreasing).  A coordinate-sorted BAM is ordered by READ position, so the
 *           second block of a read with a D / N operation starts after the first blocks of the next
 *           reads: its disorder is bounded by the reference footprint of a read, not zero (mp2_bam_open
 *           reports both numbers).  With max_span + max_disorder <= 512 the difference array is built
 *           window by window in shared memory and never touches HBM; anything else takes the general
 *           path (difference array in HBM, any order, any span).
 *           max_span = 0 = nothing is known: both numbers are measured on the device first (three small
 *           kernels, one extra pass over the events) and the path is chosen from them.
 * The call only enqueues work (the fallback is enqueued behind the fast path, gated on a device flag) unless
 * the fallback's scratch -- 4 bytes per position -- would take more than a quarter of the device's memory: then
 * the flag is read back first (one stream synchronisation) and the scratch is only allocated when it is needed.
 * Contigs that reach depth >= 511 are counted in a pool of global histograms (2048 contigs per call, depth
 * < 4607); a sample that needs more goes through the general path (exact for any depth).
 * trim_lower + trim_upper must be < 1 (the CLI clamps both, magpurify2/modules/coverage.py:39-49). */
int mp2_cov_events_device(const void *d_starts, const void *d_ends, const void *d_ev_off,
                          const void *d_contig_off, int64_t n_contigs, int64_t n_events,
                          int64_t n_positions, uint32_t max_span, uint32_t max_disorder,
                          uint32_t end_excl, float trim_lower, float trim_upper, void *d_cov,
                          int64_t cov_stride, void *d_stats, void *stream);

/* Measures, exactly, the two numbers the call above can be vouched for with (longest block; largest distance
 * of a start below the largest start before it in its contig) for events resident on the device.  One pass over
 * the events; synchronous (the results are host integers).  mp2_bam_open reports the same numbers for the BAMs it
 * decodes, so get_bam_coverages never needs this. */
int mp2_cov_order_device(const void *d_starts, const void *d_ends, const void *d_ev_off,
                         const void *d_contig_off, int64_t n_contigs, int64_t n_events,
                         uint32_t *max_span, uint32_t *max_disorder, void *stream);

/* Host-buffer form of the same call (stages H2D/D2H itself). */
int mp2_cov_events(const uint32_t *starts, const uint32_t *ends, const int64_t *ev_off,
                   const int64_t *contig_off, int64_t n_contigs, int64_t n_events,
                   uint32_t max_span, uint32_t max_disorder, uint32_t end_excl, float trim_lower,
                   float trim_upper, int device, float *cov, int64_t cov_stride, mp2_cov_stats *stats);

/* ---- scoring: replaces tools.get_log_ratio_scores / get_weighted_median ----------------- */
/* (magpurify2/tools.py:283-351; callers magpurify2/core.py:261-263 base 2, :310-312 base 25) */

/* data: float32 row-major [n_rows, n_cols] for ALL MAGs stacked; MAG m owns rows
 * mag_off[m] .. mag_off[m+1]; weights: int64[n_rows] (contig lengths).
 * col_mask: uint8[n_mags*n_cols] (1 = column takes part for that MAG) or NULL = all.
 * base / pseudo / clip_lo / clip_hi: get_log_ratio_scores' base, pseudo (1e-5), min (0), max (1).
 * min_rows: MAGs with fewer rows are not scored and get 1.0 -- 3 reproduces core.py:248-250 and
 *           :306-308 ("len(self) <= 2"), 1 is the bare function.
 * one_sided: 0 = the reference formula; 1 = only values above the median are penalised (used
 *           for the centroid-distance score, which the reference does not have).
 *   scores  : float64[n_rows]  row mean over the selected columns of
 *             1 - |ln((x+pseudo)/(median+pseudo)) / ln(base)|, clipped; MAGs with no selected
 *             column get 1.0
 *   medians : float32[n_mags*n_cols] weighted medians (NaN where not computed), or NULL        */
int mp2_logratio_device(const void *d_data, const void *d_weights, const void *d_mag_off,
                        int64_t n_mags, int64_t n_rows, int64_t n_cols, const void *d_col_mask,
                        double base, float pseudo, double clip_lo, double clip_hi,
                        int64_t min_rows, int one_sided, void *d_scores, void *d_medians,
                        void *stream);
int mp2_logratio(const float *data, const int64_t *weights, const int64_t *mag_off,
                 int64_t n_mags, int64_t n_rows, int64_t n_cols, const uint8_t *col_mask,
                 double base, float pseudo, double clip_lo, double clip_hi, int64_t min_rows,
                 int one_sided, int device, double *scores, float *medians);

/* core.Coverage sample selection (core.py:301-305): length-weighted mean per (MAG, column)
 * in f64, selected = mean >= min_average_coverage.  sel: uint8[n_mags*n_cols]. */
int mp2_select_samples_device(const void *d_cov, const void *d_weights, const void *d_mag_off,
                              int64_t n_mags, int64_t n_rows, int64_t n_cols,
                              double min_average_coverage, void *d_sel, void *d_means,
                              void *stream);

/* North-star composition score (self-specified; the reference has no CLR/centroid code):
 * clr = ln(c+1) - mean_j ln(c+1); centroid = length-weighted mean of the MAG's CLR rows;
 * dist = Euclidean distance to it.  counts: uint32[n_rows*136]; clr (float32[n_rows*136]) may be
 * NULL; dist: float64[n_rows]. */
int mp2_clr_centroid_device(const void *d_counts, const void *d_weights, const void *d_mag_off,
                            int64_t n_mags, int64_t n_rows, void *d_clr, void *d_dist,
                            void *stream);

/* ---- BAM ingest (host): replaces htslib + coverm's filter under get_bam_coverages -------- */
/* Opaque decoded sample: header names/lengths + filtered alignment blocks grouped by tid. */
typedef struct mp2_bam mp2_bam;

/* Reads a coordinate-sorted BAM with `threads` inflate workers, applies the reference's filters
 * (drop unmapped / secondary / supplementary, keep improper pairs, identity
 * 1 - NM/aligned >= min_identity in f32; src/coverage.rs:102-112) and keeps the M/=/X blocks. */
int mp2_bam_open(const char *path, float min_identity, int threads, mp2_bam **out);
int64_t mp2_bam_n_contigs(const mp2_bam *b);
int64_t mp2_bam_n_events(const mp2_bam *b);
int64_t mp2_bam_n_records(const mp2_bam *b);      /* records seen                      */
int64_t mp2_bam_n_kept(const mp2_bam *b);         /* records that passed the filters   */
uint32_t mp2_bam_max_span(const mp2_bam *b);      /* longest reference footprint of a kept read      */
uint32_t mp2_bam_max_block(const mp2_bam *b);     /* longest M/=/X block (mp2_cov_events' max_span)  */
uint32_t mp2_bam_max_disorder(const mp2_bam *b);  /* largest (block start - read position): bounds how far a
                                                     start can lie below the largest start before it       */
const char *mp2_bam_contig_name(const mp2_bam *b, int64_t i);
const int64_t *mp2_bam_contig_len(const mp2_bam *b);
const int64_t *mp2_bam_ev_off(const mp2_bam *b);
const uint32_t *mp2_bam_starts(const mp2_bam *b); /* pinned host memory */
const uint32_t *mp2_bam_ends(const mp2_bam *b);
const char *mp2_bam_header_text(const mp2_bam *b, int64_t *len);
void mp2_bam_close(mp2_bam *b);

/* ---- FASTA ingest (host): replaces tools.read_fasta under core.Mag ----------------------- */
/* (magpurify2/tools.py:210-240, core.py:35-50).  Plain or gzip.  Record id = first token of the
 * header with '~' -> '_'.  Sequence bytes are concatenated with line ends removed. */
typedef struct mp2_fasta mp2_fasta;
int mp2_fasta_open(const char *path, mp2_fasta **out);
/* Same parser on a FASTA text already in memory (bzip2 / xz inputs are decompressed by the caller). */
int mp2_fasta_open_mem(const uint8_t *data, int64_t size, mp2_fasta **out);
int64_t mp2_fasta_n(const mp2_fasta *f);
const uint8_t *mp2_fasta_bases(const mp2_fasta *f);
const int64_t *mp2_fasta_offsets(const mp2_fasta *f);
const char *mp2_fasta_id(const mp2_fasta *f, int64_t i);
const char *mp2_fasta_description(const mp2_fasta *f, int64_t i);
void mp2_fasta_close(mp2_fasta *f);

#ifdef __cplusplus
}
#endif
#endif /* MP2_H */
