// Per-MAG scoring on sm_100a: length-weighted median + log-ratio scores, sample selection,
// CLR / centroid distance.
//
// Replaces tools.get_weighted_median / tools.get_log_ratio_scores
// (/root/reference/magpurify2/tools.py:283-311, 314-351) and the glue in core.Composition /
// core.Coverage that sits directly on the native outputs (magpurify2/core.py:248-263, 301-312).
// All MAGs of a batch are stacked row-wise; MAG m owns rows mag_off[m] .. mag_off[m+1].
#include <cfloat>

#include <mutex>

#include "mp2_common.h"

namespace mp2 {

#ifndef MP2_WM_THREADS
#define MP2_WM_THREADS 256
#endif
constexpr int WM_THREADS = MP2_WM_THREADS;  // threads of a weighted-median CTA (sweep: tools/gpu_wm_sweep.sh)
constexpr int WM_SMEM_CAP = 1024;  // (value, weight) pairs sorted in shared memory (16 KB: 8 CTAs per SM); larger MAGs use scratch

struct VW {
    uint32_t key;  // order-preserving image of the f32 value (NaN last)
    float v;
    long long w;
};

__device__ __forceinline__ uint32_t f32_key(float v) {
    if (v != v) return 0xFFFFFFFFu;          // NaN sorts last (Python's sort has no defined place for it)
    if (v == 0.0f) v = 0.0f;                 // -0.0 == +0.0
    const uint32_t b = __float_as_uint(v);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}

__device__ __forceinline__ bool vw_less(const VW &a, const VW &b) {  // tuple order (value, weight), tools.py:300
    return a.key < b.key || (a.key == b.key && a.w < b.w);
}

// (value, weight) pairs of column j of the MAG's rows into a[0, n), padding that sorts last into a[n, np2)
__device__ __forceinline__ void wm_load(VW *a, const float *__restrict__ data, const long long *__restrict__ weights,
                                        int64_t r0, int64_t n, int64_t np2, int64_t n_cols, int64_t j, int tid) {
    for (int64_t i = tid; i < np2; i += WM_THREADS) {
        VW e;
        if (i < n) {
            e.v = data[(r0 + i) * n_cols + j];
            e.key = f32_key(e.v);
            e.w = weights[r0 + i];
        } else {
            e.v = 0.f; e.key = 0xFFFFFFFFu; e.w = LLONG_MAX;
        }
        a[i] = e;
    }
}

// Cumulative weights over the sorted pairs; idx = last position with 2 * cs <= total (tools.py:305-306); an exact
// tie there gives the mean of the two straddling values (:307-308), anything else the next value (:310).  The
// position and whether it is a tie travel together as 2 * idx + tie through a maximum over the CTA: the largest
// qualifying position is the one whose tie bit counts.
__device__ __forceinline__ void wm_pick(const VW *a, int64_t n, long long total, long long *s_scan, long long *s_idx,
                                        float *out, int tid) {
    const int lane = tid & 31, warp = tid >> 5;
    long long carry = 0, best = -1;  // best = largest 2 * i + tie this thread has seen
    for (int64_t b = 0; b < n; b += WM_THREADS) {
        const int64_t i = b + tid;
        const long long w = i < n ? a[i].w : 0;
        long long incl = w;
        for (int o = 1; o < 32; o <<= 1) {
            const long long t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) s_scan[warp] = incl;
        __syncthreads();
        long long wpre = 0, blk = 0;
        for (int w2 = 0; w2 < WM_THREADS / 32; w2++) {
            if (w2 < warp) wpre += s_scan[w2];
            blk += s_scan[w2];
        }
        const long long cs = carry + wpre + incl;
        if (i < n && 2 * cs <= total) best = 2 * i + (2 * cs == total ? 1 : 0);  // i grows with b
        carry += blk;
        __syncthreads();
    }
    // one value per warp, then thread 0 (a 64-bit atomicMax of every thread on one shared word is a CAS loop: it was
    // most of this kernel's time)
    for (int o = 16; o > 0; o >>= 1) {
        const long long t = __shfl_xor_sync(0xffffffffu, best, o);
        best = t > best ? t : best;
    }
    if (lane == 0) s_scan[warp] = best;
    __syncthreads();
    if (tid == 0) {
        long long packed = *s_idx;  // -1
        for (int w2 = 0; w2 < WM_THREADS / 32; w2++) packed = s_scan[w2] > packed ? s_scan[w2] : packed;
        float med;
        if (packed < 0) med = a[0].v;  // unreachable for non-negative weights (first weight <= midpoint here)
        else {
            const long long idx = packed >> 1;
            if (packed & 1) med = idx + 1 < n ? __fdiv_rn(__fadd_rn(a[idx].v, a[idx + 1].v), 2.0f) : a[idx].v;
            else med = a[idx + 1 < n ? idx + 1 : idx].v;
        }
        *out = med;
    }
}

// One CTA per (MAG, column).  medians[m * n_cols + j]; NaN for columns that are masked out or MAGs
// that are not scored (<= 2 rows).
__global__ void __launch_bounds__(WM_THREADS)
wmedian_kernel(const float *__restrict__ data, const long long *__restrict__ weights,
               const long long *__restrict__ mag_off, int64_t n_cols, const uint8_t *__restrict__ col_mask,
               int64_t min_rows, VW *__restrict__ scratch, float *__restrict__ medians) {
    extern __shared__ __align__(16) unsigned char s_raw[];
    __shared__ long long s_red[WM_THREADS / 32];
    __shared__ long long s_total, s_wmax;
    __shared__ long long s_first;
    __shared__ long long s_scan[WM_THREADS / 32];
    __shared__ long long s_idx;

    const int64_t m = blockIdx.x / n_cols, j = blockIdx.x % n_cols;
    const int64_t r0 = mag_off[m], n = mag_off[m + 1] - r0;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    float *out = medians + m * n_cols + j;
    if (n < min_rows || n < 1 || (col_mask && !col_mask[m * n_cols + j])) {
        if (tid == 0) *out = __uint_as_float(0x7FC00000u);
        return;
    }
    // ---- total weight, max weight, first row (original order) holding the max ----
    long long tot = 0, wmax = -1;
    for (int64_t i = tid; i < n; i += WM_THREADS) {
        const long long w = weights[r0 + i];
        tot += w;
        wmax = w > wmax ? w : wmax;
    }
    for (int o = 16; o > 0; o >>= 1) {
        tot += __shfl_xor_sync(0xffffffffu, tot, o);
        const long long t = __shfl_xor_sync(0xffffffffu, wmax, o);
        wmax = t > wmax ? t : wmax;
    }
    if (lane == 0) { s_red[warp] = tot; s_scan[warp] = wmax; }
    __syncthreads();
    if (tid == 0) {
        long long t = 0, x = -1;
        for (int w = 0; w < WM_THREADS / 32; w++) { t += s_red[w]; x = s_scan[w] > x ? s_scan[w] : x; }
        s_total = t; s_wmax = x; s_first = n; s_idx = -1;
    }
    __syncthreads();
    const long long total = s_total;
    if (2 * s_wmax > total) {  // tools.py:302-303: one weight above the midpoint wins
        long long f = n;
        for (int64_t i = tid; i < n; i += WM_THREADS)
            if (weights[r0 + i] == s_wmax) { f = i; break; }
        atomicMin(&s_first, f);
        __syncthreads();
        if (tid == 0) *out = data[(r0 + s_first) * n_cols + j];
        return;
    }
    // ---- sort (value, weight) ascending: bitonic network over the next power of two ----
    int64_t np2 = 1;
    while (np2 < n) np2 <<= 1;
    if (np2 <= WM_SMEM_CAP) {
        // the shared-memory pointer stays a shared-memory pointer (one pointer for both homes made every access of
        // the sort a generic LD / ST: 31 % of the kernel's stalls)
        VW *a = reinterpret_cast<VW *>(s_raw);
        wm_load(a, data, weights, r0, n, np2, n_cols, j, tid);
        __syncthreads();
        // One comparator per thread and pass (comparator t of a stage with distance s works on i = t with a zero
        // inserted at bit log2(s), and i | s), 32-bit indices.  The comparators 32b .. 32b+31 of a stage with s <= 32
        // touch the elements 64b .. 64b+63 only, and thread t always takes the comparators t, t + WM_THREADS, ...: such
        // stages follow each other with a warp barrier; only stages with s >= 64 exchange between warps.
        const int h = (int)(np2 >> 1);
        int prev_s = 64;
        for (int k = 2; k <= (int)np2; k <<= 1) {
            for (int sd = k >> 1; sd > 0; sd >>= 1) {
                if (sd >= 64 || prev_s >= 64) __syncthreads();
                else __syncwarp();
                prev_s = sd;
                for (int t = tid; t < h; t += WM_THREADS) {
                    const int i = ((t & ~(sd - 1)) << 1) | (t & (sd - 1)), p = i | sd;
                    const VW x = a[i], y = a[p];
                    const bool up = (i & k) == 0;
                    if (up ? vw_less(y, x) : vw_less(x, y)) { a[i] = y; a[p] = x; }
                }
            }
        }
        __syncthreads();
        wm_pick(a, n, total, s_scan, &s_idx, out, tid);
    } else {
        VW *a = scratch + (r0 * n_cols + j * n) * 2;
        wm_load(a, data, weights, r0, n, np2, n_cols, j, tid);
        __syncthreads();
        for (int64_t k = 2; k <= np2; k <<= 1) {
            for (int64_t s = k >> 1; s > 0; s >>= 1) {
                for (int64_t i = tid; i < np2; i += WM_THREADS) {
                    const int64_t p = i ^ s;
                    if (p > i) {
                        const VW x = a[i], y = a[p];
                        const bool up = (i & k) == 0;
                        if (up ? vw_less(y, x) : vw_less(x, y)) { a[i] = y; a[p] = x; }
                    }
                }
                __syncthreads();
            }
        }
        wm_pick(a, n, total, s_scan, &s_idx, out, tid);
    }
}

// score[r] = mean over the MAG's selected columns of 1 - |ln((x+1e-5)/(med+1e-5)) / ln(base)|, clipped
// to [0,1] (tools.py:347-351).  one_sided: only values ABOVE the median are penalised (used for the
// centroid-distance score; not a reference behaviour).
__global__ void logratio_kernel(const float *__restrict__ data, const long long *__restrict__ mag_off,
                                int64_t n_mags, int64_t n_rows, int64_t n_cols,
                                const uint8_t *__restrict__ col_mask, const float *__restrict__ medians,
                                double log_base, float pseudo, double clip_lo, double clip_hi,
                                int64_t min_rows, int one_sided, double *__restrict__ scores) {
    const int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (r >= n_rows) return;
    int64_t lo = 0, hi = n_mags;  // MAG of row r: largest m with mag_off[m] <= r
    while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if (mag_off[mid] <= r) lo = mid + 1;
        else hi = mid;
    }
    const int64_t m = lo - 1;
    const int64_t n = mag_off[m + 1] - mag_off[m];
    double acc = 0.0;
    int64_t k = 0;
    if (n >= min_rows) {
        for (int64_t j = 0; j < n_cols; j++) {
            if (col_mask && !col_mask[m * n_cols + j]) continue;
            const float num = __fadd_rn(data[r * n_cols + j], pseudo);
            const float den = __fadd_rn(medians[m * n_cols + j], pseudo);
            const float ratio = __fdiv_rn(num, den);
            const float lg = (float)log((double)ratio);  // f32 log, correctly rounded via f64
            double t = (double)lg / log_base;
            if (one_sided) t = t < 0.0 ? 0.0 : t;
            acc += 1.0 - fabs(t);
            k++;
        }
    }
    double sc;
    if (n < min_rows || k == 0) sc = 1.0;  // core.py:248-250, 306-308
    else {
        sc = k == 1 ? acc : acc / (double)k;
        if (sc < clip_lo) sc = clip_lo;  // np.clip keeps NaN
        if (sc > clip_hi) sc = clip_hi;
    }
    scores[r] = sc;
}

// core.py:301-305: np.average(cov, axis=0, weights=lengths) >= min_average_coverage, in f64, rows
// added in order (the order NumPy uses for an axis-0 reduction) so the comparison is reproducible.
__global__ void select_samples_kernel(const float *__restrict__ cov, const long long *__restrict__ weights,
                                      const long long *__restrict__ mag_off, int64_t n_mags, int64_t n_cols,
                                      double min_avg, uint8_t *__restrict__ sel, double *__restrict__ means) {
    const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t >= n_mags * n_cols) return;
    const int64_t m = t / n_cols, j = t % n_cols;
    double acc = 0.0, wsum = 0.0;
    for (int64_t r = mag_off[m]; r < mag_off[m + 1]; r++) {
        const double w = (double)weights[r];
        acc = __dadd_rn(acc, __dmul_rn((double)cov[r * n_cols + j], w));
        wsum = __dadd_rn(wsum, w);
    }
    const double mean = acc / wsum;
    if (means) means[t] = mean;
    sel[t] = mean >= min_avg ? 1 : 0;
}

// ln(c + 1) for small counts comes from a table (typical contigs hold tens of k-mers per class; the
// FP64 log was 70 % of the scoring time), larger counts are computed.
constexpr int LOG_LUT = 8192;
__device__ double g_log_lut[LOG_LUT];

__device__ __forceinline__ double ln1p_count(uint32_t c) {
    return c < LOG_LUT - 1 ? __ldg(&g_log_lut[c + 1]) : log((double)c + 1.0);
}

// The same with the first CLR_SLUT entries of the table in shared memory.  Nearly every count of a 10 kbp contig is
// below it, and 32 lanes gathering 8-byte entries from an L1-resident table cost one L1 wavefront per 128-byte line
// they touch (~15 per request): that gather, not HBM, was what clr_mag_kernel waited for.
#ifndef MP2_CLR_SLUT
#define MP2_CLR_SLUT 1024
#endif
constexpr int CLR_SLUT = MP2_CLR_SLUT;
__global__ void log_lut_kernel() {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < LOG_LUT) g_log_lut[i] = i ? log((double)i) : 0.0;
}

// Per-device, once: the table of ln(i) and the shared-memory attribute of wmedian_kernel.  Callers may be
// threads driving different GPUs (or different streams of one GPU), so this is serialised and the table is
// complete (stream synchronised) before anyone is told it exists.
static int ensure_device_state(cudaStream_t stream) {
    static std::mutex mu;
    static bool done[64] = {false};
    int dev = 0;
    MP2_CUDA(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lk(mu);
    if (dev >= 0 && dev < 64 && done[dev]) return MP2_OK;
    log_lut_kernel<<<LOG_LUT / 256, 256, 0, stream>>>();
    MP2_CUDA(cudaGetLastError());
    MP2_CUDA(cudaStreamSynchronize(stream));
    MP2_CUDA(cudaFuncSetAttribute(wmedian_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)(sizeof(VW) * WM_SMEM_CAP)));
    if (dev >= 0 && dev < 64) done[dev] = true;
    return MP2_OK;
}
static int ensure_log_lut(cudaStream_t stream) { return ensure_device_state(stream); }

// ---- CLR + centroid distance (self-specified: the reference has no such code) -----------------
// clr_ij = ln(c_ij + 1) - rm_i, rm_i = mean_j ln(c_ij + 1); centroid_j = sum_i len_i clr_ij / sum_i len_i;
// d_i = ||clr_i - centroid||_2.  f64 throughout.
//
// One CTA per MAG, two passes over the MAG's count rows (the second one hits L2), BOTH row-parallel: a warp per row,
// lane L holds the columns L, L + 32, ... (one coalesced 544-byte row per five loads), two rows in flight.
//   1. A_j = sum_i len_i ln(c_ij + 1): five running sums per lane, reduced over the CTA's warps through shared memory.
//      The row means are not needed here: sum_i len_i rm_i = mean_j A_j, so centroid_j = (A_j - mean_j A_j) / sum_i len_i.
//   2. rm_i (lane partial sums + xor butterfly), the CLR row and its distance to the centroid held in shared memory.
// ln(c + 1) comes from the shared-memory table; a count beyond it takes the out-of-line path (global table / log()),
// so that the two loops stay a few dozen instructions per row.  (The first version ran pass 1 column-parallel --
// 7 x 136 threads, each re-loading and converting the row's length -- and inlined the log() fallback into both loops:
// 90 M executed warp instructions per C3 step against 25 M now.)
#ifndef MP2_CLRM_THREADS
#define MP2_CLRM_THREADS 1024
#endif
#ifndef MP2_CLRM_MINB
#define MP2_CLRM_MINB 1
#endif
constexpr int CLRM_THREADS = MP2_CLRM_THREADS;  // sweep: tools/gpu_r2_clr_sweep.sh
constexpr int CLRM_WARPS = CLRM_THREADS / 32;
constexpr int CLRM_K = (MP2_NCANON + 31) / 32;  // columns per lane (5; the last one only on lanes 0..7)

__device__ __noinline__ double ln1p_count_slow(uint32_t c) { return ln1p_count(c); }

// this lane's CLRM_K columns of one row (ln(0 + 1) = 0: the missing columns of the last load add nothing)
__device__ __forceinline__ void clr_row_load(const uint32_t *__restrict__ row, int lane, uint32_t (&c)[CLRM_K]) {
#pragma unroll
    for (int k = 0; k < CLRM_K; k++) {
        const int j = lane + 32 * k;
        c[k] = j < MP2_NCANON ? row[j] : 0u;
    }
}
__device__ __forceinline__ void clr_logs(const uint32_t (&c)[CLRM_K], const double *s_ln, double (&v)[CLRM_K]) {
    // one warp-uniform test per row instead of a branch per count: the OR of the counts bounds their maximum
    uint32_t orr = 0;
#pragma unroll
    for (int k = 0; k < CLRM_K; k++) orr |= c[k];
    if (CLR_SLUT > 0 && !__any_sync(0xffffffffu, orr >= (uint32_t)CLR_SLUT - 1)) {
#pragma unroll
        for (int k = 0; k < CLRM_K; k++) v[k] = s_ln[c[k] + 1];
    } else {
#pragma unroll
        for (int k = 0; k < CLRM_K; k++)
            v[k] = (CLR_SLUT > 0 && c[k] < (uint32_t)CLR_SLUT - 1) ? s_ln[c[k] + 1] : ln1p_count_slow(c[k]);
    }
}
__device__ __forceinline__ void clr_row_logs(const uint32_t *__restrict__ row, int lane, const double *s_ln, double (&v)[CLRM_K]) {
    uint32_t c[CLRM_K];
    clr_row_load(row, lane, c);
    clr_logs(c, s_ln, v);
}

__global__ void __launch_bounds__(CLRM_THREADS, MP2_CLRM_MINB)
clr_mag_kernel(const uint32_t *__restrict__ counts, const long long *__restrict__ weights,
               const long long *__restrict__ mag_off, float *__restrict__ clr, double *__restrict__ dist) {
    __shared__ double s_part[CLRM_WARPS][MP2_NCANON];
    __shared__ double s_w[CLRM_WARPS];
    __shared__ double s_a[MP2_NCANON];
    __shared__ double s_cen[MP2_NCANON];
    __shared__ double s_mean_a, s_wsum;
    __shared__ double s_ln[CLR_SLUT > 0 ? CLR_SLUT : 1];
    const int64_t m = blockIdx.x;
    const int64_t r0 = mag_off[m], r1 = mag_off[m + 1];
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    for (int i = t; i < CLR_SLUT; i += CLRM_THREADS) s_ln[i] = g_log_lut[i];
    __syncthreads();
    // ---- pass 1: A_j, total length ----
    {
        double acc[CLRM_K], ws = 0.0;
#pragma unroll
        for (int k = 0; k < CLRM_K; k++) acc[k] = 0.0;
        int64_t r = r0 + warp;
        for (; r + 3 * CLRM_WARPS < r1; r += 4 * CLRM_WARPS) {  // four rows in flight: the loop waits for HBM, not for issue slots
            uint32_t c[4][CLRM_K];
            double w[4];
#pragma unroll
            for (int q = 0; q < 4; q++) {
                w[q] = (double)weights[r + q * CLRM_WARPS];
                clr_row_load(counts + (r + q * CLRM_WARPS) * MP2_NCANON, lane, c[q]);
            }
#pragma unroll
            for (int q = 0; q < 4; q++) {
                double v[CLRM_K];
                clr_logs(c[q], s_ln, v);
#pragma unroll
                for (int k = 0; k < CLRM_K; k++) acc[k] = fma(w[q], v[k], acc[k]);
                ws += w[q];  // integers below 2^53: exact in any order
            }
        }
        for (; r < r1; r += CLRM_WARPS) {
            double va[CLRM_K];
            const double wa = (double)weights[r];
            clr_row_logs(counts + r * MP2_NCANON, lane, s_ln, va);
#pragma unroll
            for (int k = 0; k < CLRM_K; k++) acc[k] = fma(wa, va[k], acc[k]);
            ws += wa;
        }
#pragma unroll
        for (int k = 0; k < CLRM_K; k++) {
            const int j = lane + 32 * k;
            if (j < MP2_NCANON) s_part[warp][j] = acc[k];
        }
        if (lane == 0) s_w[warp] = ws;
    }
    __syncthreads();
    if (t < MP2_NCANON) {
        double a = 0.0;
        for (int w = 0; w < CLRM_WARPS; w++) a += s_part[w][t];
        s_a[t] = a;
    }
    __syncthreads();
    if (warp == 0) {
        double sa = 0.0;
        for (int j = lane; j < MP2_NCANON; j += 32) sa += s_a[j];
        for (int o = 16; o > 0; o >>= 1) sa += __shfl_xor_sync(0xffffffffu, sa, o);
        if (lane == 0) {
            double w = 0.0;
            for (int k = 0; k < CLRM_WARPS; k++) w += s_w[k];
            s_mean_a = sa / (double)MP2_NCANON;
            s_wsum = w;
        }
    }
    __syncthreads();
    if (t < MP2_NCANON) s_cen[t] = s_wsum > 0.0 ? (s_a[t] - s_mean_a) / s_wsum : 0.0;
    __syncthreads();
    // ---- pass 2: row mean, CLR row, distance to the centroid ----
    double cen[CLRM_K];
#pragma unroll
    for (int k = 0; k < CLRM_K; k++) cen[k] = lane + 32 * k < MP2_NCANON ? s_cen[lane + 32 * k] : 0.0;
    // Four FP64 operations per count here (sum, minus centroid, minus row mean, fused square-add) and one in pass 1.
    // (Taking the distance from sum_j u_j^2 - 136 rm_i^2 needs three and ONE butterfly -- 0.128 instead of 0.140 ms,
    // the kernel waits on its dependent chains, not on the FP64 pipe (25 % busy) -- but its cancellation turns the exact 0
    // of a contig that IS the centroid into 1e-6.)
    auto finish_row = [&](int64_t r, const double (&v)[CLRM_K]) {
        double s = 0.0, u[CLRM_K];
#pragma unroll
        for (int k = 0; k < CLRM_K; k++) {
            u[k] = v[k] - cen[k];  // missing columns: 0 - 0
            s += v[k];
        }
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        const double rm = s / (double)MP2_NCANON;
        double d2 = 0.0;
#pragma unroll
        for (int k = 0; k < CLRM_K; k++) {
            const int j = lane + 32 * k;
            if (j < MP2_NCANON) {
                if (clr) clr[r * MP2_NCANON + j] = (float)(v[k] - rm);
                const double d = u[k] - rm;
                d2 = fma(d, d, d2);
            }
        }
        for (int o = 16; o > 0; o >>= 1) d2 += __shfl_xor_sync(0xffffffffu, d2, o);
        if (lane == 0) dist[r] = sqrt(d2);
    };
    int64_t r = r0 + warp;
    for (; r + 3 * CLRM_WARPS < r1; r += 4 * CLRM_WARPS) {  // four rows in flight (L2 latency)
        uint32_t c[4][CLRM_K];
#pragma unroll
        for (int q = 0; q < 4; q++) clr_row_load(counts + (r + q * CLRM_WARPS) * MP2_NCANON, lane, c[q]);
#pragma unroll
        for (int q = 0; q < 4; q++) {
            double v[CLRM_K];
            clr_logs(c[q], s_ln, v);
            finish_row(r + q * CLRM_WARPS, v);
        }
    }
    for (; r < r1; r += CLRM_WARPS) {
        double v[CLRM_K];
        clr_row_logs(counts + r * MP2_NCANON, lane, s_ln, v);
        finish_row(r, v);
    }
}

}  // namespace mp2

using namespace mp2;

extern "C" int mp2_logratio_device(const void *d_data, const void *d_weights, const void *d_mag_off,
                                   int64_t n_mags, int64_t n_rows, int64_t n_cols, const void *d_col_mask,
                                   double base, float pseudo, double clip_lo, double clip_hi,
                                   int64_t min_rows, int one_sided, void *d_scores, void *d_medians,
                                   void *stream_) {
    if (n_mags < 0 || n_rows < 0 || n_cols < 1) return fail(MP2_EINVAL, "bad shape");
    if (!(base > 0.0) || base == 1.0) return fail(MP2_EINVAL, "base must be > 0 and != 1");
    if (!d_data || !d_weights || !d_mag_off || !d_scores) return fail(MP2_EINVAL, "NULL argument");
    if (n_mags * n_cols > INT32_MAX) return fail(MP2_EINVAL, "too many (MAG, column) pairs");
    MP2_TRY(require_device(-1));
    cudaStream_t stream = (cudaStream_t)stream_;
    if (n_rows == 0 || n_mags == 0) return MP2_OK;
    Temp med, scratch;
    float *medians = (float *)d_medians;
    if (!medians) {
        MP2_TRY(med.alloc(sizeof(float) * n_mags * n_cols, stream));
        medians = med.as<float>();
    }
    // scratch for MAGs with more rows than fit in shared memory: 2x (power-of-two padding) pairs
    MP2_TRY(scratch.alloc(sizeof(VW) * 2 * (size_t)n_rows * n_cols, stream));
    MP2_TRY(ensure_device_state(stream));
    {
        Launch L("wmedian_kernel", stream);
        wmedian_kernel<<<(unsigned)(n_mags * n_cols), WM_THREADS, sizeof(VW) * WM_SMEM_CAP, stream>>>(
            (const float *)d_data, (const long long *)d_weights, (const long long *)d_mag_off, n_cols,
            (const uint8_t *)d_col_mask, min_rows, scratch.as<VW>(), medians);
    }
    MP2_CUDA(cudaGetLastError());
    {
        Launch L("logratio_kernel", stream);
        logratio_kernel<<<(unsigned)((n_rows + 255) / 256), 256, 0, stream>>>(
            (const float *)d_data, (const long long *)d_mag_off, n_mags, n_rows, n_cols,
            (const uint8_t *)d_col_mask, medians, log(base), pseudo, clip_lo, clip_hi, min_rows, one_sided,
            (double *)d_scores);
    }
    MP2_CUDA(cudaGetLastError());
    return MP2_OK;
}

extern "C" int mp2_select_samples_device(const void *d_cov, const void *d_weights, const void *d_mag_off,
                                         int64_t n_mags, int64_t n_rows, int64_t n_cols,
                                         double min_average_coverage, void *d_sel, void *d_means,
                                         void *stream_) {
    if (n_mags < 0 || n_rows < 0 || n_cols < 1) return fail(MP2_EINVAL, "bad shape");
    if (!d_cov || !d_weights || !d_mag_off || !d_sel) return fail(MP2_EINVAL, "NULL argument");
    MP2_TRY(require_device(-1));
    cudaStream_t stream = (cudaStream_t)stream_;
    if (n_mags == 0) return MP2_OK;
    const int64_t t = n_mags * n_cols;
    {
        Launch L("select_samples_kernel", stream);
        select_samples_kernel<<<(unsigned)((t + 127) / 128), 128, 0, stream>>>(
            (const float *)d_cov, (const long long *)d_weights, (const long long *)d_mag_off, n_mags, n_cols,
            min_average_coverage, (uint8_t *)d_sel, (double *)d_means);
    }
    MP2_CUDA(cudaGetLastError());
    return MP2_OK;
}

extern "C" int mp2_clr_centroid_device(const void *d_counts, const void *d_weights, const void *d_mag_off,
                                       int64_t n_mags, int64_t n_rows, void *d_clr, void *d_dist,
                                       void *stream_) {
    if (n_mags < 0 || n_rows < 0) return fail(MP2_EINVAL, "bad shape");
    if (!d_counts || !d_weights || !d_mag_off || !d_dist) return fail(MP2_EINVAL, "NULL argument");
    MP2_TRY(require_device(-1));
    cudaStream_t stream = (cudaStream_t)stream_;
    if (n_rows == 0 || n_mags == 0) return MP2_OK;
    MP2_TRY(ensure_log_lut(stream));
    (void)n_rows;
    {
        Launch L("clr_mag_kernel", stream);
        clr_mag_kernel<<<(unsigned)n_mags, CLRM_THREADS, 0, stream>>>(
            (const uint32_t *)d_counts, (const long long *)d_weights, (const long long *)d_mag_off, (float *)d_clr,
            (double *)d_dist);
    }
    MP2_CUDA(cudaGetLastError());
    return MP2_OK;
}
