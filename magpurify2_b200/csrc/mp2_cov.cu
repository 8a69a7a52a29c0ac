// Coverage path: per-contig depth accumulation + trimmed-mean coverage for ONE sample on sm_100a.
//
// Replaces the arithmetic coverm 0.3.2 performs under _coverage.get_bam_coverages
// (/root/reference/src/coverage.rs:99-142: contig_coverage -> TrimmedMean add_contig /
// calculate_coverage; restated in SURVEY.md Appendix A because the crate is not vendored).
//
// Layout in HBM: one int32 difference array D over ALL contigs concatenated (contig c owns
// positions off[c] .. off[c+1]) plus one trailing slot.  An alignment block [s, e) of contig c
// adds +1 at off[c]+s and -1 at off[c]+min(e, len): the reference skips the -1 when e == len
// (`if end < len`), here it lands on the first slot of the NEXT contig, which makes the plain
// (unsegmented) prefix sum of D equal to the per-contig depth everywhere -- no segment heads.
//
//   cov_scatter_kernel   events -> D (red.global +-1) and per-tile sums of D (tile = 65536 slots)
//   cov_tile_scan_kernel exclusive scan of the tile sums  => depth just before every tile
//   cov_depth_kernel     ONE pass over D: running depth, windowed depth histogram per contig in
//                        shared memory (red.shared: equal depths in a warp are merged by the
//                        hardware), trimmed mean for contigs whose window lies in one tile;
//                        windows spanning tiles are merged through a global row and finished by
//                        the last contributor.  D is read exactly once (4 B/position).
//   cov_deep_kernel      exact fallback for contigs that reach depth >= 1023 (multi-pass over
//                        depth ranges; rare).
#include <algorithm>
#include <climits>
#include <cstddef>
#include <cstdlib>
#include <mutex>

#include "mp2_common.h"

namespace mp2 {

constexpr int COV_THREADS = 256;
constexpr int COV_ITEMS = 16;                         // positions per thread per sub-tile
constexpr int COV_SUB = COV_THREADS * COV_ITEMS;      // 4096 positions
constexpr int COV_TILE = 65536;                       // positions per tile
constexpr int COV_NSUB = COV_TILE / COV_SUB;          // 16
constexpr int COV_BINS = 512;                         // depth bins 0..510 exact, 511 = overflow (2 KB of shared
                                                      // memory per warp: 16 warps per SM in cov_stream_kernel; 1024
                                                      // bins: 14 warps, 4.5 % slower on C3)
constexpr int COV_EV_TILE = 2048;                     // events per scatter CTA
constexpr int COV_DEEP_BINS = 8192;

// row bookkeeping for windows that span tiles (one row per tile: the window that STARTS in it)
struct RowMeta {
    unsigned long long done;  // window positions accounted for so far
    unsigned int maxd;        // highest depth seen (1023 = overflow)
    unsigned int pad;
};

// ---- small device helpers ------------------------------------------------------------------
__device__ __forceinline__ void smem_inc(uint32_t saddr) {
    asm volatile("red.shared.add.u32 [%0], 1;" ::"r"(saddr) : "memory");
}

__device__ __forceinline__ void smem_add(uint32_t saddr, int v) {
    asm volatile("red.shared.add.s32 [%0], %1;" ::"r"(saddr), "r"(v) : "memory");
}
// A shared-window address the compiler cannot see through: a plain cvta result is rematerialised at every use
// (S2UR SR_CgaCtaId + UMOV + ULEA, and the ULEA waits on the S2UR), this one stays in a register.
__device__ __forceinline__ uint32_t smem_base_opaque(const void *p) {
    uint32_t r = (uint32_t)__cvta_generic_to_shared(p);
    asm volatile("mov.u32 %0, %0;" : "+r"(r));
    return r;
}

// predicated form: no branch around the reduction
__device__ __forceinline__ void smem_inc_if(uint32_t saddr, uint32_t cond) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %1, 0;\n\t@p red.shared.add.u32 [%0], 1;\n\t}" ::"r"(saddr), "r"(cond) : "memory");
}

__device__ __forceinline__ void smem_dec_if(uint32_t saddr, uint32_t cond) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %1, 0;\n\t@p red.shared.add.s32 [%0], -1;\n\t}" ::"r"(saddr), "r"(cond) : "memory");
}

// upper_bound(a[lo..hi), x) - 1 : largest i in [lo, hi) with a[i] <= x (lo if none)
__device__ __forceinline__ int64_t last_le(const int64_t *__restrict__ a, int64_t lo, int64_t hi, int64_t x) {
    int64_t l = lo, h = hi;
    while (l < h) {
        int64_t m = (l + h) >> 1;
        if (a[m] <= x) l = m + 1;
        else h = m;
    }
    return l > lo ? l - 1 : lo;
}

struct TrimParams {
    float lo;   // trim_lower
    float hi;   // 1.0f - trim_upper   (src/coverage.rs:99)
    uint32_t excl;
};

// Appendix A.4 (calculate_coverage) for one contig, executed by ONE warp.  bins[d] = number of
// window positions with depth d, d in [0, nbins).  The sequential histogram walk of the reference
// is evaluated in closed form per bin from the running count A_d:
//   d0 = first d with A_d >= min_index            contributes (min(A_d, max_index) - min_index + 1) * d
//   d > d0                                        contributes max(0, min(A_d, max_index+1) - A_{d-1}) * d
// `acc`/`total` carry the state across calls (depth ranges of the fallback kernel).
template <class Load>
__device__ __forceinline__ void trim_walk_warp(Load load, uint32_t d_begin, uint32_t d_end, uint64_t min_index,
                                               uint64_t max_index, unsigned long long &acc,
                                               unsigned long long &total) {
    const int lane = threadIdx.x & 31;
    for (uint32_t d0 = d_begin; d0 < d_end; d0 += 32) {
        const uint32_t d = d0 + lane;
        unsigned long long cnt = d < d_end ? (unsigned long long)load(d) : 0ull;
        unsigned long long incl = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            unsigned long long t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        const unsigned long long A = acc + incl, Ap = A - cnt;
        unsigned long long take = 0;
        if (d < d_end) {
            const bool first = (A >= min_index) && (d == 0 || Ap < min_index);
            if (first) take = (A > max_index ? max_index : A) - min_index + 1;
            else if (d > 0 && Ap >= min_index) {
                const unsigned long long top = A < max_index + 1 ? A : max_index + 1;
                take = top > Ap ? top - Ap : 0;
            }
        }
        unsigned long long part = take * d;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
        total += part;
        acc = __shfl_sync(0xffffffffu, A, 31);
    }
}

__device__ __forceinline__ void trim_indices(uint64_t N, const TrimParams &tp, uint64_t &mn, uint64_t &mx) {
    const float fn = __ull2float_rn(N);
    mn = (uint64_t)floorf(__fmul_rn(tp.lo, fn));
    mx = (uint64_t)ceilf(__fmul_rn(tp.hi, fn));
}

__device__ __forceinline__ void write_result(int64_t c, uint64_t N, uint64_t covered, uint64_t mn, uint64_t mx,
                                             uint64_t total, uint32_t maxd, uint32_t flags,
                                             float *__restrict__ cov, int64_t cov_stride,
                                             mp2_cov_stats *__restrict__ stats) {
    float r = 0.0f;
    if (N != 0 && covered != 0) r = __fdiv_rn(__ull2float_rn(total), __ull2float_rn(mx - mn));
    cov[c * cov_stride] = r;
    if (stats) {
        mp2_cov_stats s;
        s.observed_len = N; s.covered = covered; s.min_index = mn; s.max_index = mx;
        s.total = covered ? total : 0; s.max_depth = maxd; s.coverage = r; s.flags = flags;
        stats[c] = s;
    }
}

// ---- per-contig initialisation: contigs without a window get their final answer here --------
__global__ void cov_init_kernel(const int64_t *__restrict__ off, int64_t n, uint32_t excl,
                                float *__restrict__ cov, int64_t cov_stride,
                                mp2_cov_stats *__restrict__ stats, int *__restrict__ gate) {
    int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (c >= n) return;
    const int64_t len = off[c + 1] - off[c];
    if (len >= (1ll << 30)) *gate = 1;  // the tile path keeps 32-bit tile-relative coordinates
    if (2ll * excl >= len) write_result(c, 0, 0, 0, 0, 0, 0, 0, cov, cov_stride, stats);
}

// ---- tile -> first contig maps ------------------------------------------------------------------
__global__ void cov_map_kernel(const int64_t *__restrict__ a, int64_t n, int64_t step, int64_t n_tiles,
                               int32_t *__restrict__ first) {
    int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (k >= n_tiles) return;
    first[k] = (int32_t)last_le(a, 0, n, k * step);  // a[n] (the total) is never a candidate
}

// ---- scatter -----------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
cov_scatter_kernel(const uint32_t *__restrict__ starts, const uint32_t *__restrict__ ends,
                   const int64_t *__restrict__ ev_off, const int64_t *__restrict__ off, int64_t n,
                   int64_t n_events, const int32_t *__restrict__ ev_first, int32_t *__restrict__ D,
                   int32_t *__restrict__ tile_sum, const int *__restrict__ gate) {
    if (gate && *gate == 0) return;  // fallback leg with nothing to do
    for (int64_t tile = blockIdx.x; tile * COV_EV_TILE < n_events; tile += gridDim.x) {
    const int64_t t0 = tile * COV_EV_TILE;
    const int64_t c_lo = ev_first[tile];
    const int64_t t_last = t0 + COV_EV_TILE - 1 < n_events - 1 ? t0 + COV_EV_TILE - 1 : n_events - 1;
    __shared__ int64_t s_chi;
    if (threadIdx.x == 0) s_chi = last_le(ev_off, c_lo, n, t_last);
    __syncthreads();
    const int64_t c_hi = s_chi;
#pragma unroll
    for (int k = 0; k < COV_EV_TILE / 256; k++) {
        const int64_t e = t0 + k * 256 + threadIdx.x;
        const bool live = e < n_events;
        int64_t ps = -1, pe = -1;
        if (live) {
            const int64_t c = c_hi == c_lo ? c_lo : last_le(ev_off, c_lo, c_hi + 1, e);
            const int64_t base = off[c], len = off[c + 1] - base;
            const int64_t s = starts[e];
            int64_t en = ends[e];
            if (en > len) en = len;
            if (s < len && en > s) {  // a start beyond the contig would panic in the reference
                ps = base + s;
                pe = base + en;
                atomicAdd(&D[ps], 1);
                atomicAdd(&D[pe], -1);
            }
        }
        // tile sums of D: a block whose +1 and -1 fall in the same tile contributes nothing
        if (ps >= 0) {
            const int64_t ts = ps / COV_TILE, te = pe / COV_TILE;
            if (ts != te) {
                atomicAdd(&tile_sum[ts], 1);
                atomicAdd(&tile_sum[te], -1);
            }
        }
    }
    __syncthreads();  // s_chi is rewritten by the next tile
    }
}

// gated zero fill of D (general path used as a fallback: skip the 4 B/slot write when not needed)
__global__ void cov_zero_kernel(int4 *__restrict__ p, int64_t n16, const int *__restrict__ gate) {
    if (gate && *gate == 0) return;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n16; i += stride) p[i] = make_int4(0, 0, 0, 0);
}

// ---- exclusive scan of the tile sums (n_tiles is small: L / 65536) -------------------------------
__global__ void __launch_bounds__(1024)
cov_tile_scan_kernel(const int32_t *__restrict__ tile_sum, int64_t n_tiles, int32_t *__restrict__ tile_pre,
                     const int *__restrict__ gate) {
    if (gate && *gate == 0) return;
    __shared__ int s_w[32];
    __shared__ int s_carry;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) s_carry = 0;
    __syncthreads();
    for (int64_t b = 0; b < n_tiles; b += 1024) {
        const int64_t i = b + tid;
        const int v = i < n_tiles ? tile_sum[i] : 0;
        int incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) s_w[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            int w = s_w[lane], wi = w;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                int t = __shfl_up_sync(0xffffffffu, wi, o);
                if (lane >= o) wi += t;
            }
            s_w[lane] = wi - w;  // exclusive
        }
        __syncthreads();
        const int carry = s_carry;
        if (i < n_tiles) tile_pre[i] = carry + s_w[warp] + incl - v;
        __syncthreads();
        if (tid == 1023) s_carry = carry + s_w[warp] + incl;
        __syncthreads();
    }
}

// ---- fused scan + windowed depth histogram + trimmed mean -----------------------------------------
struct DepthArgs {
    const int32_t *D;
    const int64_t *off;
    int64_t n, n_pos;  // n_pos = off[n] + 1 slots in D
    const int32_t *first;     // tile -> contig holding the tile's first position
    const int32_t *tile_pre;  // depth just before the tile
    int64_t n_tiles;
    uint32_t *rows;           // [n_tiles][COV_BINS]
    RowMeta *meta;            // [n_tiles]
    int32_t *deep_list;       // contigs needing the exact fallback
    int *deep_count;
    int *work_counter;
    const int *gate;          // general path as a fallback: run only when *gate != 0 (NULL = always)
    float *cov;
    int64_t cov_stride;
    mp2_cov_stats *stats;
    TrimParams tp;
};

// Finish contig c from a histogram (shared or global), by warp 0 of the CTA.
template <class Load>
__device__ __forceinline__ void finish_contig(const DepthArgs &a, int64_t c, uint64_t N, uint32_t maxd, Load load) {
    const int lane = threadIdx.x & 31;
    if (maxd >= COV_BINS - 1) {  // overflow: exact multi-pass fallback
        if (lane == 0) a.deep_list[atomicAdd(a.deep_count, 1)] = (int32_t)c;
        return;
    }
    uint64_t mn, mx;
    trim_indices(N, a.tp, mn, mx);
    const uint64_t covered = N - (uint64_t)load(0);
    unsigned long long acc = 0, total = 0;
    if (covered) trim_walk_warp(load, 0, maxd + 1, mn, mx, acc, total);
    if (lane == 0) write_result(c, N, covered, mn, mx, total, maxd, MP2_COV_GENERAL_PATH, a.cov, a.cov_stride, a.stats);
}

__global__ void __launch_bounds__(COV_THREADS)
cov_depth_kernel(const DepthArgs a) {
    if (a.gate && *a.gate == 0) return;
    __shared__ uint32_t s_hist[COV_BINS];
    __shared__ int s_wsum[COV_THREADS / 32];
    __shared__ unsigned int s_maxd;
    __shared__ int s_tile;
    __shared__ int s_last;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t h0 = (uint32_t)__cvta_generic_to_shared(s_hist);
    const int64_t excl = a.tp.excl;
    for (int i = tid; i < COV_BINS; i += COV_THREADS) s_hist[i] = 0;
    if (tid == 0) s_maxd = 0;
    __syncthreads();

    for (;;) {
        if (tid == 0) s_tile = atomicAdd(a.work_counter, 1);
        __syncthreads();
        const int64_t tile = s_tile;
        __syncthreads();
        if (tile >= a.n_tiles) break;
        const int64_t tile_lo = tile * COV_TILE;
        const int64_t tile_hi = tile_lo + COV_TILE < a.n_pos ? tile_lo + COV_TILE : a.n_pos;
        int carry = a.tile_pre[tile];

        // current contig: the first one whose window can still receive positions >= tile_lo
        int64_t c = a.first[tile];
        int64_t wlo = 0, whi = -1;  // window [wlo, whi] of contig c (global positions), empty if none
        bool have = false;          // window of c loaded
        uint64_t in_tile = 0;       // window positions of c histogrammed in this tile so far

        for (int sub = 0; sub < COV_NSUB; sub++) {
            const int64_t p0 = tile_lo + (int64_t)sub * COV_SUB;
            if (p0 >= tile_hi) break;
            const int64_t p1 = p0 + COV_SUB < tile_hi ? p0 + COV_SUB : tile_hi;
            // ---- load 16 slots per thread, local + block scan ----
            const int64_t q = p0 + (int64_t)tid * COV_ITEMS;
            int d[COV_ITEMS];
            if (q + COV_ITEMS <= a.n_pos && ((reinterpret_cast<uintptr_t>(a.D + q) & 15) == 0)) {
                const int4 *src = reinterpret_cast<const int4 *>(a.D + q);
#pragma unroll
                for (int k = 0; k < COV_ITEMS / 4; k++) {
                    const int4 v = __ldcs(src + k);
                    d[4 * k] = v.x; d[4 * k + 1] = v.y; d[4 * k + 2] = v.z; d[4 * k + 3] = v.w;
                }
            } else {
#pragma unroll
                for (int k = 0; k < COV_ITEMS; k++) d[k] = q + k < a.n_pos ? a.D[q + k] : 0;
            }
#pragma unroll
            for (int k = 1; k < COV_ITEMS; k++) d[k] += d[k - 1];
            int incl = d[COV_ITEMS - 1];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                int t = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += t;
            }
            if (lane == 31) s_wsum[warp] = incl;
            __syncthreads();
            int wpre = 0, sub_total = 0;
#pragma unroll
            for (int w = 0; w < COV_THREADS / 32; w++) {
                const int v = s_wsum[w];
                if (w < warp) wpre += v;
                sub_total += v;
            }
            const int tpre = carry + wpre + incl - d[COV_ITEMS - 1];  // depth before this thread's first slot
#pragma unroll
            for (int k = 0; k < COV_ITEMS; k++) d[k] += tpre;         // d[k] = depth at position q + k
            carry += sub_total;

            // ---- windows intersecting [p0, p1) ----
            for (;;) {
                if (!have) {
                    if (c >= a.n) break;
                    const int64_t o0 = a.off[c], o1 = a.off[c + 1];
                    if (o0 >= p1) break;  // contig starts after this sub-tile
                    if (o1 - o0 > 2 * excl) { wlo = o0 + excl; whi = o1 - excl - 1; }
                    else { wlo = 0; whi = -1; }
                    have = true;
                    in_tile = 0;
                }
                if (whi < wlo || whi < tile_lo) { c++; have = false; continue; }  // no window / ended before this tile
                if (wlo >= p1) break;
                // positions [lo, hi) of this sub-tile belong to the window
                const int64_t lo = wlo > p0 ? wlo : p0, hi = whi + 1 < p1 ? whi + 1 : p1;
                if (hi > lo) {
                    const int ja = lo - q > 0 ? (lo - q > COV_ITEMS ? COV_ITEMS : (int)(lo - q)) : 0;
                    const int jb = hi - q < COV_ITEMS ? (hi - q < 0 ? 0 : (int)(hi - q)) : COV_ITEMS;
                    int mx = 0;
#pragma unroll
                    for (int k = 0; k < COV_ITEMS; k++) {
                        if (k >= ja && k < jb) {
                            int v = d[k] < 0 ? 0 : d[k];
                            v = v > COV_BINS - 1 ? COV_BINS - 1 : v;
                            mx = v > mx ? v : mx;
                            smem_inc(h0 + 4u * (uint32_t)v);
                        }
                    }
                    mx = __reduce_max_sync(0xffffffffu, mx);
                    if (lane == 0 && mx) atomicMax(&s_maxd, (unsigned)mx);
                    in_tile += (uint64_t)(hi - lo);
                }
                if (whi >= p1) break;  // window continues in the next sub-tile
                // ---- the window of contig c ends in this sub-tile ----
                __syncthreads();
                const uint64_t N = (uint64_t)(whi - wlo + 1);
                const uint32_t maxd = s_maxd;
                if (wlo >= tile_lo) {
                    // whole window inside this tile: finish from shared memory
                    if (warp == 0) finish_contig(a, c, N, maxd, [&](uint32_t i) { return s_hist[i]; });
                } else {
                    // window started in an earlier tile: merge into its row, last contributor finishes
                    const int64_t row = wlo / COV_TILE;
                    uint32_t *r = a.rows + row * COV_BINS;
                    for (uint32_t i = tid; i <= maxd; i += COV_THREADS) {
                        const uint32_t v = s_hist[i];
                        if (v) atomicAdd(&r[i], v);
                    }
                    __threadfence();
                    __syncthreads();
                    if (tid == 0) {
                        atomicMax(&a.meta[row].maxd, maxd);
                        __threadfence();
                        const unsigned long long old = atomicAdd(&a.meta[row].done, (unsigned long long)in_tile);
                        s_last = (old + in_tile == N);
                    }
                    __syncthreads();
                    if (s_last) {
                        __threadfence();
                        const uint32_t gm = *((volatile unsigned int *)&a.meta[row].maxd);
                        if (warp == 0) finish_contig(a, c, N, gm, [&](uint32_t i) { return __ldcg(r + i); });
                    }
                }
                __syncthreads();
                for (uint32_t i = tid; i <= maxd; i += COV_THREADS) s_hist[i] = 0;
                if (tid == 0) s_maxd = 0;
                __syncthreads();
                c++;
                have = false;
            }
            __syncthreads();  // s_wsum is rewritten by the next sub-tile
        }
        // ---- tile end: a window still open is merged into the row of the tile it started in ----
        if (have && whi >= wlo && in_tile > 0) {
            __syncthreads();
            const uint64_t N = (uint64_t)(whi - wlo + 1);
            const uint32_t maxd = s_maxd;
            const int64_t row = wlo / COV_TILE;
            uint32_t *r = a.rows + row * COV_BINS;
            for (uint32_t i = tid; i <= maxd; i += COV_THREADS) {
                const uint32_t v = s_hist[i];
                if (v) atomicAdd(&r[i], v);
            }
            __threadfence();
            __syncthreads();
            if (tid == 0) {
                atomicMax(&a.meta[row].maxd, maxd);
                __threadfence();
                const unsigned long long old = atomicAdd(&a.meta[row].done, (unsigned long long)in_tile);
                s_last = (old + in_tile == N);
            }
            __syncthreads();
            if (s_last) {
                __threadfence();
                const uint32_t gm = *((volatile unsigned int *)&a.meta[row].maxd);
                if (warp == 0) finish_contig(a, c, N, gm, [&](uint32_t i) { return __ldcg(r + i); });
            }
            __syncthreads();
            for (uint32_t i = tid; i <= maxd; i += COV_THREADS) s_hist[i] = 0;
            if (tid == 0) s_maxd = 0;
            __syncthreads();
        }
    }
}

// ---- exact fallback for very deep contigs ------------------------------------------------------
// One CTA per flagged contig; the depth axis is swept in ranges of COV_DEEP_BINS bins, each range
// re-scanning the contig's slice of D (the prefix restarts at 0 at a contig start).
__global__ void __launch_bounds__(COV_THREADS)
cov_deep_kernel(const DepthArgs a, int *__restrict__ deep_next) {
    if (a.gate && *a.gate == 0) return;
    __shared__ uint32_t s_hist[COV_DEEP_BINS];
    __shared__ int s_wsum[COV_THREADS / 32];
    __shared__ unsigned int s_maxd;
    __shared__ int s_item;
    __shared__ unsigned long long s_acc, s_total;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t excl = a.tp.excl;
    for (;;) {
        if (tid == 0) s_item = atomicAdd(deep_next, 1);
        __syncthreads();
        const int item = s_item;
        __syncthreads();
        if (item >= *a.deep_count) break;
        const int64_t c = a.deep_list[item];
        const int64_t o0 = a.off[c], o1 = a.off[c + 1];
        const int64_t wlo = o0 + excl, whi = o1 - excl - 1;
        const uint64_t N = (uint64_t)(whi - wlo + 1);
        uint64_t mn, mx;
        trim_indices(N, a.tp, mn, mx);
        if (tid == 0) { s_maxd = 0; s_acc = 0; s_total = 0; }
        uint64_t covered = 0;
        uint32_t maxd = 0;
        for (uint32_t base = 0;; base += COV_DEEP_BINS) {
            for (int i = tid; i < COV_DEEP_BINS; i += COV_THREADS) s_hist[i] = 0;
            __syncthreads();
            int carry = 0;
            for (int64_t p0 = o0; p0 < o1; p0 += COV_SUB) {
                const int64_t q = p0 + (int64_t)tid * COV_ITEMS;
                int d[COV_ITEMS];
#pragma unroll
                for (int k = 0; k < COV_ITEMS; k++) d[k] = q + k < o1 ? a.D[q + k] : 0;
#pragma unroll
                for (int k = 1; k < COV_ITEMS; k++) d[k] += d[k - 1];
                int incl = d[COV_ITEMS - 1];
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    int t = __shfl_up_sync(0xffffffffu, incl, o);
                    if (lane >= o) incl += t;
                }
                if (lane == 31) s_wsum[warp] = incl;
                __syncthreads();
                int wpre = 0, sub_total = 0;
#pragma unroll
                for (int w = 0; w < COV_THREADS / 32; w++) {
                    const int v = s_wsum[w];
                    if (w < warp) wpre += v;
                    sub_total += v;
                }
                const int tpre = carry + wpre + incl - d[COV_ITEMS - 1];
                carry += sub_total;
                unsigned int mxl = 0;
#pragma unroll
                for (int k = 0; k < COV_ITEMS; k++) {
                    const int64_t p = q + k;
                    if (p >= wlo && p <= whi) {
                        const int v = d[k] + tpre;
                        const uint32_t u = v < 0 ? 0u : (uint32_t)v;
                        mxl = u > mxl ? u : mxl;
                        if (u >= base && u - base < COV_DEEP_BINS) atomicAdd(&s_hist[u - base], 1u);
                    }
                }
                mxl = __reduce_max_sync(0xffffffffu, mxl);
                if (lane == 0 && base == 0) atomicMax(&s_maxd, mxl);
                __syncthreads();
            }
            maxd = s_maxd;
            if (base == 0) covered = N - s_hist[0];
            if (warp == 0) {
                unsigned long long acc = s_acc, total = s_total;
                const uint32_t d_end = maxd + 1 < base + COV_DEEP_BINS ? maxd + 1 : base + COV_DEEP_BINS;
                if (covered) trim_walk_warp([&](uint32_t i) { return s_hist[i - base]; }, base, d_end, mn, mx, acc, total);
                if (lane == 0) { s_acc = acc; s_total = total; }
            }
            __syncthreads();
            if ((uint64_t)base + COV_DEEP_BINS > maxd) break;
        }
        if (tid == 0) write_result(c, N, covered, mn, mx, s_total, maxd, MP2_COV_GENERAL_PATH, a.cov, a.cov_stride, a.stats);
        __syncthreads();
    }
}


// =====================================================================================================
// Fast path: the difference array never leaves shared memory.
//
// Precondition (checked by cov_order_kernel unless the caller vouches for it with max_span > 0):
// starts are non-decreasing inside every contig and end - start <= max_span.  Then the blocks that
// can touch the slots [lo, hi) of the concatenated coordinate space are one contiguous range of
// the event arrays (global start in [lo - max_span, hi)), found by two binary searches.
// One CTA owns a tile of 16384 slots: zero a 68 KB shared array, apply the range's +1 / -1 with
// shared-memory reductions, count the blocks that are open at the tile start (= the depth just
// before the tile), then run the same scan / window histogram / trimmed-mean code as the
// general path, reading the differences from shared memory.  HBM traffic: the events, once
// (plus max_span/16384 of them twice), and the results.
// =====================================================================================================
constexpr int CT_TILE = 16384;
constexpr int CT_NSUB = CT_TILE / COV_SUB;        // 4
constexpr int CT_WORDS = CT_TILE + CT_TILE / 16;  // one pad word per 16: a thread's 16 slots are conflict-free
constexpr size_t CT_SMEM = sizeof(int32_t) * CT_WORDS;

__device__ __forceinline__ int ct_phys(int i) { return i + (i >> 4); }

// first event whose global start (off[contig] + start) is >= X
__device__ __forceinline__ int64_t first_event_ge(const uint32_t *__restrict__ starts, const int64_t *__restrict__ ev_off,
                                                  const int64_t *__restrict__ off, int64_t n, int64_t X) {
    if (X <= 0) return 0;
    if (X >= off[n]) return ev_off[n];
    const int64_t c = last_le(off, 0, n, X);  // contig holding slot X
    const int64_t want = X - off[c];
    int64_t lo = ev_off[c], hi = ev_off[c + 1];
    while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if ((int64_t)starts[mid] < want) lo = mid + 1;
        else hi = mid;
    }
    return lo;
}

__global__ void cov_bounds_kernel(const uint32_t *__restrict__ starts, const int64_t *__restrict__ ev_off,
                                  const int64_t *__restrict__ off, int64_t n, int64_t n_tiles, uint32_t max_span,
                                  const unsigned int *__restrict__ d_span, int64_t *__restrict__ e_lo,
                                  int64_t *__restrict__ e_hi, int32_t *__restrict__ c_lo,
                                  int32_t *__restrict__ c_hi) {
    const int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (k >= n_tiles) return;
    if (d_span) max_span = d_span[1];  // measured by cov_order_kernel
    const int64_t A = k * CT_TILE - (int64_t)max_span, B = (k + 1) * CT_TILE;
    e_lo[k] = first_event_ge(starts, ev_off, off, n, A);
    e_hi[k] = first_event_ge(starts, ev_off, off, n, B);
    // every event of [e_lo, e_hi) belongs to a contig overlapping the slots [A, B)
    c_lo[k] = (int32_t)last_le(off, 0, n, A > 0 ? A : 0);
    c_hi[k] = (int32_t)last_le(off, 0, n, B - 1 < off[n] ? B - 1 : off[n]);
}

// ---- order / span measurement (max_span == 0: nothing is vouched for) ---------------------------------
// The fast paths need two numbers: the longest block (end - start) and the DISORDER of the starts,
//   disorder = max_j ( max_{i<j} g_i - g_j ),   g = contig offset + start  (monotone across contigs),
// i.e. how far a start may fall behind the running maximum.  A BAM is sorted by read position, so the first
// blocks of the reads are in order, but the second block of a read with a D / N operation starts after the
// next reads' first blocks: the disorder is bounded by the reference footprint of a read (a few hundred bp),
// not zero.  flag[0] = 1: starts go back (old kernels need disorder == 0), flag[1] = longest
// block, flag[2] = disorder.  Events that are dropped anyway (start beyond the contig, empty) do not count.
// One pass over the events: every CTA takes a tile of 2048 consecutive events, every thread 8 consecutive ones
// (a contig cursor walks along, so the contig of an event costs a comparison, not a search).  Per tile:
// largest and smallest key, the disorder INSIDE the tile (against the tile's own running maximum) and the
// longest block.  The disorder against everything before the tile is (largest key before the tile) - (smallest
// key of the tile); cov_order_scan_kernel takes that maximum over the tiles.  Together: exact.
struct OrderTile {
    long long kmax, kmin;
};

__global__ void __launch_bounds__(256)
cov_order_kernel(const uint32_t *__restrict__ starts, const uint32_t *__restrict__ ends,
                 const int64_t *__restrict__ ev_off, const int64_t *__restrict__ off, int64_t n, int64_t n_events,
                 const int32_t *__restrict__ ev_first, OrderTile *__restrict__ tiles, unsigned int *__restrict__ flag) {
    constexpr int PER = COV_EV_TILE / 256;  // consecutive events per thread
    __shared__ long long s_w[8];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t t0 = (int64_t)blockIdx.x * COV_EV_TILE;
    const int64_t e0 = t0 + (int64_t)tid * PER;
    // A thread's eight events nearly always lie in ONE contig (a contig holds ~2000 of them): their keys are then
    // kept relative to the contig, 32 bits, with 0xFFFFFFFF for an event that does not count; only a thread that
    // sees a contig boundary (or the end of the events) keeps 64-bit keys.  (All keys in 64 bits: 91 instructions per
    // event, 1.66 G per C3 sample, the ALU pipe 88 % busy: 2.0 ms for one pass over 4.7 GB.)
    long long g[PER];
    uint32_t key[PER];
    long long base = 0;
    bool rel = false;
    long long tmax = -1, tmin = LLONG_MAX;
    unsigned int span = 0;
    if (e0 < n_events) {
        // contigs of this tile: ev_first[tile] .. ev_first[tile + 1] (the map has one entry beyond the last tile)
        const int64_t c_hi = (int64_t)ev_first[blockIdx.x + 1] + 1;
        int64_t c = last_le(ev_off, ev_first[blockIdx.x], c_hi < n ? c_hi : n, e0);
        int64_t next = ev_off[c + 1], len;
        base = off[c];
        len = off[c + 1] - base;
        uint32_t sv[PER], ev[PER];
        if (e0 + PER <= n_events && ((reinterpret_cast<uintptr_t>(starts + e0) | reinterpret_cast<uintptr_t>(ends + e0)) & 15) == 0) {
#pragma unroll
            for (int k = 0; k < PER; k += 4) {
                const uint4 a4 = __ldg(reinterpret_cast<const uint4 *>(starts + e0 + k));
                const uint4 b4 = __ldg(reinterpret_cast<const uint4 *>(ends + e0 + k));
                sv[k] = a4.x; sv[k + 1] = a4.y; sv[k + 2] = a4.z; sv[k + 3] = a4.w;
                ev[k] = b4.x; ev[k + 1] = b4.y; ev[k + 2] = b4.z; ev[k + 3] = b4.w;
            }
        } else {
#pragma unroll
            for (int k = 0; k < PER; k++) {
                sv[k] = e0 + k < n_events ? __ldg(starts + e0 + k) : 0u;
                ev[k] = e0 + k < n_events ? __ldg(ends + e0 + k) : 0u;
            }
        }
        rel = e0 + PER <= next && e0 + PER <= n_events;
        if (rel) {
            const uint32_t len32 = len > 0xFFFFFFFFll ? 0xFFFFFFFFu : (uint32_t)len;
            uint32_t kmax = 0, kmin = 0xFFFFFFFFu;
#pragma unroll
            for (int k = 0; k < PER; k++) {
                const uint32_t en = ev[k] < len32 ? ev[k] : len32;
                const bool ok = sv[k] < en;  // start inside the contig, block not empty: anything else is dropped by every path
                key[k] = ok ? sv[k] : 0xFFFFFFFFu;
                const uint32_t sp = ok ? en - sv[k] : 0u;
                span = sp > span ? sp : span;
                kmax = ok && sv[k] > kmax ? sv[k] : kmax;
                kmin = key[k] < kmin ? key[k] : kmin;
            }
            if (kmin != 0xFFFFFFFFu) { tmax = base + kmax; tmin = base + kmin; }
        } else {
#pragma unroll
            for (int k = 0; k < PER; k++) {
                const int64_t e = e0 + k;
                g[k] = -1;
                if (e < n_events) {
                    while (e >= next) {  // next contig that has events
                        c++;
                        next = ev_off[c + 1];
                        base = off[c];
                        len = off[c + 1] - base;
                    }
                    const int64_t s_ = sv[k];
                    int64_t en = ev[k];
                    if (en > len) en = len;
                    if (s_ < len && en > s_) {
                        g[k] = base + s_;
                        const unsigned int sp = (unsigned int)(en - s_);
                        span = sp > span ? sp : span;
                        tmax = g[k] > tmax ? g[k] : tmax;
                        tmin = g[k] < tmin ? g[k] : tmin;
                    }
                }
            }
        }
    } else {
#pragma unroll
        for (int k = 0; k < PER; k++) g[k] = -1;
    }
    // exclusive prefix maximum of the thread maxima inside the tile
    long long incl = tmax;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const long long t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o && t > incl) incl = t;
    }
    long long pm = __shfl_up_sync(0xffffffffu, incl, 1);
    if (lane == 0) pm = -1;
    if (lane == 31) s_w[warp] = incl;
    __syncthreads();
    for (int w = 0; w < warp; w++) pm = s_w[w] > pm ? s_w[w] : pm;
    unsigned long long dis = 0;
    if (rel) {
        // the running maximum relative to this contig: a maximum that lies in an earlier contig (pm < base) is below
        // every key of this one, like 0
        uint32_t cur = pm > base ? (uint32_t)(pm - base) : 0u, d = 0;
#pragma unroll
        for (int k = 0; k < PER; k++) {
            if (key[k] != 0xFFFFFFFFu) {
                const uint32_t back = cur > key[k] ? cur - key[k] : 0u;
                d = back > d ? back : d;
                cur = key[k] > cur ? key[k] : cur;
            }
        }
        dis = d;
    } else {
#pragma unroll
        for (int k = 0; k < PER; k++) {
            if (g[k] >= 0) {
                if (pm > g[k] && (unsigned long long)(pm - g[k]) > dis) dis = (unsigned long long)(pm - g[k]);
                pm = g[k] > pm ? g[k] : pm;
            }
        }
    }
    unsigned int d32 = dis > 0xFFFFFFFFull ? 0xFFFFFFFFu : (unsigned int)dis;
    span = __reduce_max_sync(0xffffffffu, span);
    d32 = __reduce_max_sync(0xffffffffu, d32);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const long long t = __shfl_xor_sync(0xffffffffu, tmin, o);
        tmin = t < tmin ? t : tmin;
    }
    __shared__ long long s_min[8];
    if (lane == 0) {
        s_min[warp] = tmin;
        // the three results are running maxima: a warp that cannot raise them does not touch them (one atomic per
        // warp on the same word was 4.5 of the kernel's 6.5 ms on a sample whose starts are out of order everywhere)
        if (d32) {
            if (!(__ldcg(&flag[0]) & 1u)) atomicOr(&flag[0], 1u);
            if (d32 > __ldcg(&flag[2])) atomicMax(&flag[2], d32);
        }
        if (span > __ldcg(&flag[1])) atomicMax(&flag[1], span);
    }
    __syncthreads();
    if (tid == 0) {
        long long mx = -1, mn = LLONG_MAX;
        for (int w = 0; w < 8; w++) {
            mx = s_w[w] > mx ? s_w[w] : mx;
            mn = s_min[w] < mn ? s_min[w] : mn;
        }
        tiles[blockIdx.x].kmax = mx;
        tiles[blockIdx.x].kmin = mn;
    }
}

// disorder across tiles: max over the tiles of (largest key before the tile) - (smallest key of the tile).
// One CTA, 8 consecutive tiles per thread.
__global__ void __launch_bounds__(1024)
cov_order_scan_kernel(const OrderTile *__restrict__ tiles, int64_t n_tiles, unsigned int *__restrict__ flag) {
    constexpr int PER = 8;
    __shared__ long long s_w[32];
    __shared__ long long s_carry;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) s_carry = -1;
    __syncthreads();
    unsigned long long dis = 0;
    for (int64_t b = 0; b < n_tiles; b += 1024 * PER) {
        const int64_t i0 = b + (int64_t)tid * PER;
        long long mx[PER], mn[PER];
        long long tmax = -1;
#pragma unroll
        for (int k = 0; k < PER; k++) {
            const bool in = i0 + k < n_tiles;
            mx[k] = in ? tiles[i0 + k].kmax : -1;
            mn[k] = in ? tiles[i0 + k].kmin : LLONG_MAX;
            tmax = mx[k] > tmax ? mx[k] : tmax;
        }
        long long incl = tmax;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const long long t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o && t > incl) incl = t;
        }
        long long pre = __shfl_up_sync(0xffffffffu, incl, 1);
        if (lane == 0) pre = -1;
        if (lane == 31) s_w[warp] = incl;
        __syncthreads();
        const long long carry = s_carry;
        pre = carry > pre ? carry : pre;
        for (int w = 0; w < warp; w++) pre = s_w[w] > pre ? s_w[w] : pre;  // largest key before this thread's tiles
#pragma unroll
        for (int k = 0; k < PER; k++) {
            if (mn[k] != LLONG_MAX && pre > mn[k] && (unsigned long long)(pre - mn[k]) > dis)
                dis = (unsigned long long)(pre - mn[k]);
            pre = mx[k] > pre ? mx[k] : pre;
        }
        __syncthreads();
        if (tid == 1023) s_carry = pre;
        __syncthreads();
    }
    unsigned int d32 = dis > 0xFFFFFFFFull ? 0xFFFFFFFFu : (unsigned int)dis;
    d32 = __reduce_max_sync(0xffffffffu, d32);
    if (lane == 0 && d32) { atomicOr(&flag[0], 1u); atomicMax(&flag[2], d32); }
}

constexpr int CT_CTAB = 96;  // contigs of a tile whose offsets are staged in shared memory
// Deep pool: contigs whose depth leaves the shared histogram claim a row of it.  A dominant organism sequenced
// beyond 500x brings ALL its contigs here at once (hundreds), so the pool is wide rather than tall: 2048 rows of
// 4096 bins (32 MB, cleared per sample).  Beyond that -- more deep contigs, or depth >= 4606 -- the sample is
// redone by the general path (2.5x slower, same results).  (256 rows x 32768 bins sent 2 of 80 C3-shaped
// samples of an 8-GPU run down the general path: one rank at 87 ms per step, the others at 58.)
constexpr int CT_DEEP_ROWS = 2048;    // fewest rows of the pool = contigs per sample that may leave the shared histogram's depth
                                      // range; the pool grows with the number of contigs of the call (deep_pool_rows)
constexpr int CT_DEEP_BINS = 4096;    // depths COV_BINS-1 .. COV_BINS-2+4096 are counted in the deep pool

struct TileArgs {
    const uint32_t *starts, *ends;
    const int64_t *ev_off, *off;
    int64_t n, n_pos, n_tiles;
    const int64_t *e_lo, *e_hi;
    const int32_t *c_lo, *c_hi;  // contigs the tile's events can belong to
    const int32_t *c_ev;         // cov_stream2_kernel: a contig at or before the one owning the run's first event
    int64_t n_events;            // all events of the sample (bulk copies never read beyond their 16-byte granule)
    uint32_t *rows;              // [n_tiles][COV_BINS]
    RowMeta *meta;
    int32_t *deep_list;
    int *deep_count;
    int *work_counter;
    const int32_t *order;       // cov_stream2_kernel: runs in the order they are drawn (most events first), or NULL
    int32_t *deep_row_of;       // [n] -1 = none; row of the deep pool claimed by the contig
    uint32_t *deep_pool;        // [CT_DEEP_ROWS][CT_DEEP_BINS] counts of depths >= COV_BINS - 1
    int *deep_rows_used;
    int deep_rows;              // rows of deep_pool
    const unsigned int *check;  // NULL, or check[0] != 0 => the precondition failed: do nothing
    uint32_t span_host;         // max_span the bounds were computed with (vouched by the caller) ...
    const unsigned int *span_dev;  // ... or measured on the device (then span_host is ignored)
    int *gate_out;              // set to 1 when the general path has to run instead
    float *cov;
    int64_t cov_stride;
    mp2_cov_stats *stats;
    TrimParams tp;
};

// Finish contig c by warp 0.  Depths 0..COV_BINS-2 come from `load` (shared histogram or the
// contig's global row); deeper positions were counted in the contig's row of the deep pool.
template <class Load>
__device__ __forceinline__ void finish_contig_t(const TileArgs &a, int64_t c, uint64_t N, uint32_t maxd, Load load) {
    const int lane = threadIdx.x & 31;
    const uint32_t *deep = nullptr;
    if (maxd >= COV_BINS - 1) {
        const int r = *((volatile int32_t *)&a.deep_row_of[c]);
        if (r < 0 || maxd >= (uint32_t)(COV_BINS - 1 + CT_DEEP_BINS)) {  // pool exhausted / beyond its range
            if (lane == 0) {
                a.deep_list[atomicAdd(a.deep_count, 1)] = (int32_t)c;
                *a.gate_out = 1;  // the exact general path redoes this sample
            }
            return;
        }
        deep = a.deep_pool + (size_t)r * CT_DEEP_BINS;
    }
    uint64_t mn, mx;
    trim_indices(N, a.tp, mn, mx);
    const uint64_t covered = N - (uint64_t)load(0);
    unsigned long long acc = 0, total = 0;
    if (covered) {
        const uint32_t top = maxd + 1 < COV_BINS - 1 ? maxd + 1 : COV_BINS - 1;
        trim_walk_warp(load, 0, top, mn, mx, acc, total);
        if (deep)
            trim_walk_warp([&](uint32_t i) { return __ldcg(deep + (i - (COV_BINS - 1))); }, COV_BINS - 1, maxd + 1, mn, mx,
                           acc, total);
    }
    if (lane == 0) write_result(c, N, covered, mn, mx, total, maxd, 0, a.cov, a.cov_stride, a.stats);
}

// Rare: depths beyond the shared histogram's range are counted in the contig's row of the deep pool
// (claimed on first use).  Called by the lanes of a warp that hold such positions.
__device__ __forceinline__ void ct_deep_positions(const TileArgs &a, int64_t c, const int *d, uint32_t m) {
    const int lane = threadIdx.x & 31;
    int r = 0;
    const unsigned grp = __activemask();
    const int leader = __ffs(grp) - 1;
    if (lane == leader) {
        r = atomicAdd(&a.deep_row_of[c], 0);
        if (r == -1) {
            const int old = atomicCAS(&a.deep_row_of[c], -1, -2);
            if (old == -1) {
                const int k2 = atomicAdd(a.deep_rows_used, 1);
                r = k2 < a.deep_rows ? k2 : -3;
                atomicExch(&a.deep_row_of[c], r);
            } else {
                r = old;
            }
        }
        while (r == -2) r = atomicAdd(&a.deep_row_of[c], 0);
    }
    r = __shfl_sync(grp, r, leader);
    if (r >= 0) {
        uint32_t *dr = a.deep_pool + (size_t)r * CT_DEEP_BINS;
        for (int k = 0; k < COV_ITEMS; k++) {
            const int v = d[k] - (COV_BINS - 1);
            if ((m & (1u << k)) && v >= 0 && v < CT_DEEP_BINS) atomicAdd(&dr[v], 1u);
        }
        __threadfence();
    }
}

__device__ __forceinline__ int clamp_rel(int64_t v) {  // tile-relative coordinate, saturated
    return v < -(1ll << 30) ? -(1 << 30) : (v > (1ll << 30) ? (1 << 30) : (int)v);
}

// All coordinates inside the kernel are 32-bit and relative to the tile start (or to the tile's
// first event); the contigs a tile can see are staged in shared memory (offset, length, first
// event), CT_CTAB at a time -- a tile crowded with more (tiny) contigs walks them in batches.
__global__ void __launch_bounds__(COV_THREADS, 3)
cov_tile_kernel(const TileArgs a) {
    extern __shared__ __align__(16) int32_t s_T[];  // CT_WORDS
    __shared__ uint32_t s_hist[COV_BINS];
    __shared__ int s_wsum[COV_THREADS / 32];
    __shared__ unsigned int s_maxd;
    __shared__ int s_tile, s_last, s_carry;
    __shared__ int s_crel[CT_CTAB + 1];  // contig start - tile_lo (saturated)
    __shared__ int s_cev[CT_CTAB + 1];   // contig's first event - E0 (saturated)

    if (a.check && a.check[0] != 0) {
        if (blockIdx.x == 0 && threadIdx.x == 0) *a.gate_out = 1;
        return;
    }
    if (*((volatile int *)a.gate_out)) return;  // the general path redoes the whole sample anyway
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t h0 = (uint32_t)__cvta_generic_to_shared(s_hist);
    const int excl = (int)a.tp.excl;
    for (int i = tid; i < COV_BINS; i += COV_THREADS) s_hist[i] = 0;
    if (tid == 0) s_maxd = 0;
    __syncthreads();

    for (;;) {
        if (tid == 0) s_tile = atomicAdd(a.work_counter, 1);
        __syncthreads();
        const int64_t tile = s_tile;
        if (tile >= a.n_tiles) break;
        const int64_t tile_lo = tile * CT_TILE;
        const int tlen = (int)(tile_lo + CT_TILE < a.n_pos ? CT_TILE : a.n_pos - tile_lo);
        const int64_t E0 = a.e_lo[tile], E1 = a.e_hi[tile];
        const int nev = (int)(E1 - E0);
        const int64_t ca = a.c_lo[tile], cb = a.c_hi[tile];

        // ---- zero the tile, stage the contig table ----
        for (int i = tid; i < CT_WORDS / 4; i += COV_THREADS) reinterpret_cast<int4 *>(s_T)[i] = make_int4(0, 0, 0, 0);
        if (tid == 0) s_carry = 0;
        int open = 0;
        // ---- events: +1 at the start, -1 at the end, blocks open across the tile start ----
        for (int64_t cbase = ca; cbase <= cb; cbase += CT_CTAB) {
            const int nc = (int)(cb - cbase + 1 < CT_CTAB ? cb - cbase + 1 : CT_CTAB);
            __syncthreads();
            for (int i = tid; i <= nc; i += COV_THREADS) {
                s_crel[i] = clamp_rel(a.off[cbase + i] - tile_lo);
                s_cev[i] = clamp_rel(a.ev_off[cbase + i] - E0);
            }
            __syncthreads();
            // events of this batch of contigs: [max(0, cev[0]), min(nev, cev[nc]))
            const int eb0 = s_cev[0] > 0 ? s_cev[0] : 0, eb1 = s_cev[nc] < nev ? s_cev[nc] : nev;
            const int cv1 = nc > 1 ? s_cev[1] : INT_MAX, cv2 = nc > 2 ? s_cev[2] : INT_MAX, cv3 = nc > 3 ? s_cev[3] : INT_MAX;
            for (int e0 = eb0 + tid; e0 < eb1; e0 += 4 * COV_THREADS) {
                uint32_t sv[4], ev[4];
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    const int e = e0 + u * COV_THREADS;
                    sv[u] = e < eb1 ? __ldg(a.starts + E0 + e) : 0u;
                    ev[u] = e < eb1 ? __ldg(a.ends + E0 + e) : 0u;
                }
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    const int e = e0 + u * COV_THREADS;
                    if (e >= eb1) break;
                    int j = 0;
                    if (nc <= 4) {
                        j = (e >= cv1) + (e >= cv2) + (e >= cv3);
                    } else {
                        int lo = 0, hi = nc;  // largest j in [0, nc) with s_cev[j] <= e
                        while (lo < hi) {
                            const int mid = (lo + hi) >> 1;
                            if (s_cev[mid] <= e) lo = mid + 1;
                            else hi = mid;
                        }
                        j = lo > 0 ? lo - 1 : 0;
                    }
                    const int crel = s_crel[j];
                    const uint32_t len = (uint32_t)(s_crel[j + 1] - crel);  // exact unless saturated (then irrelevant)
                    const uint32_t s = sv[u];
                    const uint32_t en = ev[u] < len ? ev[u] : len;
                    if (s < len && en > s) {
                        const uint32_t gs = (uint32_t)crel + s, ge = (uint32_t)crel + en;  // mod 2^32
                        const uint32_t ncrel = crel < 0 ? (uint32_t)(-crel) : 0u;
                        if (s >= ncrel) {
                            if (gs < (uint32_t)tlen) atomicAdd(&s_T[ct_phys((int)gs)], 1);
                        } else if (en >= ncrel) {
                            open++;  // started before the tile and still covers slot tile_lo - 1
                        }
                        if (en >= ncrel && ge < (uint32_t)tlen) atomicAdd(&s_T[ct_phys((int)ge)], -1);
                    }
                }
            }
        }
        open = __reduce_add_sync(0xffffffffu, open);
        if (lane == 0 && open) atomicAdd(&s_carry, open);
        __syncthreads();
        int carry = s_carry;

        // ---- scan + windows; contigs intersecting the tile are a sub-range of [ca, cb] ----
        int64_t cbase = ca;                       // contig table currently staged: [cbase, cbase + nct]
        int nct = (int)(cb - cbase + 1 < CT_CTAB ? cb - cbase + 1 : CT_CTAB);
        if (cb - ca + 1 > CT_CTAB) {              // the event loop left the LAST batch staged: restage the first
            __syncthreads();
            for (int i = tid; i <= nct; i += COV_THREADS) s_crel[i] = clamp_rel(a.off[cbase + i] - tile_lo);
            __syncthreads();
        }
        int64_t c = ca;
        int wlo = 0, whi = -1;
        bool have = false;
        int in_tile = 0;

        for (int sub = 0; sub < CT_NSUB; sub++) {
            const int p0 = sub * COV_SUB;
            if (p0 >= tlen) break;
            const int p1 = p0 + COV_SUB < tlen ? p0 + COV_SUB : tlen;
            const int q = p0 + tid * COV_ITEMS;
            int d[COV_ITEMS];
            {
                const int b = ct_phys(q);
#pragma unroll
                for (int k = 0; k < COV_ITEMS; k++) d[k] = s_T[b + k];
            }
#pragma unroll
            for (int k = 1; k < COV_ITEMS; k++) d[k] += d[k - 1];
            int incl = d[COV_ITEMS - 1];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                int t = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += t;
            }
            if (lane == 31) s_wsum[warp] = incl;
            __syncthreads();
            int wpre = 0, sub_total = 0;
#pragma unroll
            for (int w = 0; w < COV_THREADS / 32; w++) {
                const int v = s_wsum[w];
                if (w < warp) wpre += v;
                sub_total += v;
            }
            const int tpre = carry + wpre + incl - d[COV_ITEMS - 1];
#pragma unroll
            for (int k = 0; k < COV_ITEMS; k++) d[k] += tpre;
            carry += sub_total;

            for (;;) {
                if (!have) {
                    if (c > cb) break;
                    if (c - cbase >= nct) {  // next batch of the contig table
                        __syncthreads();
                        cbase = c;
                        nct = (int)(cb - cbase + 1 < CT_CTAB ? cb - cbase + 1 : CT_CTAB);
                        for (int i = tid; i <= nct; i += COV_THREADS) s_crel[i] = clamp_rel(a.off[cbase + i] - tile_lo);
                        __syncthreads();
                    }
                    const int o0 = s_crel[c - cbase], o1 = s_crel[c - cbase + 1];
                    if (o0 >= p1) break;
                    if (o1 - o0 > 2 * excl) { wlo = o0 + excl; whi = o1 - excl - 1; }  // saturated ends stay far outside
                    else { wlo = 0; whi = -1; }
                    have = true;
                    in_tile = 0;
                }
                if (whi < wlo || whi < 0) { c++; have = false; continue; }
                if (wlo >= p1) break;
                const int lo = wlo > p0 ? wlo : p0, hi = whi + 1 < p1 ? whi + 1 : p1;
                if (hi > lo) {
                    const int ja = lo - q > 0 ? (lo - q > COV_ITEMS ? COV_ITEMS : lo - q) : 0;
                    const int jb = hi - q < COV_ITEMS ? (hi - q < 0 ? 0 : hi - q) : COV_ITEMS;
                    if (jb > ja) {
                        const uint32_t m = ((1u << jb) - 1u) & ~((1u << ja) - 1u);
                        int mx = 0;
                        if (m == 0xFFFFu) {  // all 16 slots of this thread are inside the window (the common case)
#pragma unroll
                            for (int k = 0; k < COV_ITEMS; k++) mx = d[k] > mx ? d[k] : mx;
                            if (mx < COV_BINS - 1) {
#pragma unroll
                                for (int k = 0; k < COV_ITEMS; k++) smem_inc(h0 + 4u * (uint32_t)d[k]);
                            } else {
#pragma unroll
                                for (int k = 0; k < COV_ITEMS; k++)
                                    if (d[k] < COV_BINS - 1) smem_inc(h0 + 4u * (uint32_t)d[k]);
                            }
                        } else {
#pragma unroll
                            for (int k = 0; k < COV_ITEMS; k++) {
                                const int v = d[k];
                                const uint32_t in = m & (1u << k);
                                mx = (in && v > mx) ? v : mx;
                                if (in && v < COV_BINS - 1) smem_inc(h0 + 4u * (uint32_t)v);
                            }
                        }
                        if (mx >= COV_BINS - 1) {  // rare; the copy keeps d[] itself in registers
                            int dd[COV_ITEMS];
#pragma unroll
                            for (int k = 0; k < COV_ITEMS; k++) dd[k] = d[k];
                            ct_deep_positions(a, c, dd, m);
                        }
                        mx = __reduce_max_sync(__activemask(), mx);
                        if (mx > (int)s_maxd) atomicMax(&s_maxd, (unsigned)mx);
                    }
                    in_tile += hi - lo;
                }
                if (whi >= p1) break;
                // ---- the window of contig c ends in this sub-tile ----
                __syncthreads();
                const int64_t o0g = a.off[c], o1g = a.off[c + 1];
                const uint64_t N = (uint64_t)(o1g - o0g - 2 * (int64_t)excl);
                const uint32_t maxd = s_maxd;
                const uint32_t mtop = maxd < COV_BINS - 1 ? maxd : COV_BINS - 1;
                if (wlo >= 0) {
                    if (warp == 0) finish_contig_t(a, c, N, maxd, [&](uint32_t i) { return s_hist[i]; });
                } else {
                    const int64_t row = (o0g + excl) / CT_TILE;
                    uint32_t *r = a.rows + row * COV_BINS;
                    for (uint32_t i = tid; i <= mtop; i += COV_THREADS) {
                        const uint32_t v = s_hist[i];
                        if (v) atomicAdd(&r[i], v);
                    }
                    __threadfence();
                    __syncthreads();
                    if (tid == 0) {
                        atomicMax(&a.meta[row].maxd, maxd);
                        __threadfence();
                        const unsigned long long old = atomicAdd(&a.meta[row].done, (unsigned long long)in_tile);
                        s_last = (old + (unsigned long long)in_tile == N);
                    }
                    __syncthreads();
                    if (s_last) {
                        __threadfence();
                        const uint32_t gm = *((volatile unsigned int *)&a.meta[row].maxd);
                        if (warp == 0) finish_contig_t(a, c, N, gm, [&](uint32_t i) { return __ldcg(r + i); });
                    }
                }
                __syncthreads();
                for (uint32_t i = tid; i <= mtop; i += COV_THREADS) s_hist[i] = 0;
                if (tid == 0) s_maxd = 0;
                __syncthreads();
                c++;
                have = false;
            }
            __syncthreads();
        }
        if (have && whi >= wlo && in_tile > 0) {
            __syncthreads();
            const int64_t o0g = a.off[c], o1g = a.off[c + 1];
            const uint64_t N = (uint64_t)(o1g - o0g - 2 * (int64_t)excl);
            const uint32_t maxd = s_maxd;
                const uint32_t mtop = maxd < COV_BINS - 1 ? maxd : COV_BINS - 1;
            const int64_t row = (o0g + excl) / CT_TILE;
            uint32_t *r = a.rows + row * COV_BINS;
            for (uint32_t i = tid; i <= mtop; i += COV_THREADS) {
                const uint32_t v = s_hist[i];
                if (v) atomicAdd(&r[i], v);
            }
            __threadfence();
            __syncthreads();
            if (tid == 0) {
                atomicMax(&a.meta[row].maxd, maxd);
                __threadfence();
                const unsigned long long old = atomicAdd(&a.meta[row].done, (unsigned long long)in_tile);
                s_last = (old + (unsigned long long)in_tile == N);
            }
            __syncthreads();
            if (s_last) {
                __threadfence();
                const uint32_t gm = *((volatile unsigned int *)&a.meta[row].maxd);
                if (warp == 0) finish_contig_t(a, c, N, gm, [&](uint32_t i) { return __ldcg(r + i); });
            }
            __syncthreads();
            for (uint32_t i = tid; i <= mtop; i += COV_THREADS) s_hist[i] = 0;
            if (tid == 0) s_maxd = 0;
            __syncthreads();
        }
        __syncthreads();  // s_T / s_tile are rewritten by the next tile
    }
}


// =====================================================================================================
// cov_stream_kernel: the tile algorithm with NO block barrier.  A warp (= CTA) owns a run of 65536
// slots and streams through it in windows of 2048 slots held in its own shared memory: events are
// consumed in order, 32 per batch (one per lane); an end that falls beyond the current window lands
// in a 512-slot overflow region that becomes the head of the next window, so the running depth is
// simply carried from window to window (only the first window of a run looks back max_span slots).
// The contig histogram stays in the warp's shared memory across windows until the contig ends; only
// contigs crossing a RUN boundary go through the global rows.  Needs max_span <= 512.
// =====================================================================================================
#ifndef MP2_C3_SPL
#define MP2_C3_SPL 64
#endif
constexpr int C3_SPL = MP2_C3_SPL;           // consecutive slots per lane (a multiple of 16)
constexpr int C3_SPL_LOG = C3_SPL == 64 ? 6 : 5;
static_assert(C3_SPL == 64 || C3_SPL == 32, "slots per lane");
constexpr int C3_W = 32 * C3_SPL;            // slots per window
constexpr int C3_S = 512;                    // overflow region = largest block span of the fast path
constexpr int C3_RUN = 65536;                // slots per run
constexpr int C3_TPR = C3_RUN / C3_W;        // windows per run
constexpr int C3_WORDS = (C3_W + C3_S) + 4 * ((C3_W + C3_S) / C3_SPL);  // 4 pad words per lane's worth of slots
constexpr int C3_HEAD = C3_S + 4 * (C3_S / C3_SPL);                     // words of the overflow region
#ifndef MP2_C3_PF
#define MP2_C3_PF 384
#endif
constexpr int C3_PF = MP2_C3_PF;             // events ahead of e_cur that are prefetched into L2

__device__ __forceinline__ int c3_phys(int i) { return i + 4 * (i >> C3_SPL_LOG); }

// weight class of a run with n_ev events: 0 = heaviest; a class is a factor sqrt(2) wide
constexpr int C4_WEIGHT_CLASSES = 64;
__device__ __forceinline__ int run_weight_class(int64_t n_ev) {
    uint32_t v = n_ev < 0 ? 1u : n_ev >= 0x7fffffff ? 0x80000000u : (uint32_t)n_ev + 1u;
    const int lz = 31 - __clz(v);                       // floor(log2 v)
    const int half = lz ? (int)((v >> (lz - 1)) & 1u) : 0;
    return C4_WEIGHT_CLASSES - 1 - (2 * lz + half);    // 2 * 31 + 1 = 63 at most
}

// Event range and contigs of every run.  `meas` = {unsorted, longest block, disorder} measured by
// cov_order_kernel, or NULL when the caller vouched for (max_span, max_disorder).  The ranges are computed with
// the LARGEST margins the kernels support (span + disorder <= C3_S), so they do not depend on the measured
// values: with starts that fall at most `disorder` behind their running maximum, a binary search for X ends
// between the first event whose running maximum reaches X and the first one whose running maximum reaches
// X + disorder.  Hence every event before search(A - C3_S) ends at or before slot A, and every event from
// search(B + C3_S) on starts at or after slot B.
__global__ void cov_bounds_run_kernel(const uint32_t *__restrict__ starts, const int64_t *__restrict__ ev_off,
                                      const int64_t *__restrict__ off, int64_t n, int64_t n_runs, uint32_t max_span,
                                      uint32_t max_disorder, int need_sorted, const unsigned int *__restrict__ meas,
                                      int64_t *__restrict__ e_lo, int64_t *__restrict__ e_hi,
                                      int32_t *__restrict__ c_lo, int32_t *__restrict__ c_hi,
                                      int32_t *__restrict__ c_ev, int *__restrict__ gate,
                                      unsigned int *__restrict__ weight_count) {
    const int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (k >= n_runs) return;
    if (meas) { max_span = meas[1]; max_disorder = meas[2]; }
    if ((uint64_t)max_span + max_disorder > (uint64_t)C3_S || (need_sorted && max_disorder)) {
        if (k == 0) *gate = 1;  // blocks too long / starts too far out of order for the overflow region: general path
        return;
    }
    const int64_t A = k * C3_RUN - (int64_t)C3_S, B = (k + 1) * C3_RUN;
    e_lo[k] = first_event_ge(starts, ev_off, off, n, A);
    e_hi[k] = first_event_ge(starts, ev_off, off, n, need_sorted ? B : B + C3_S);
    c_lo[k] = (int32_t)last_le(off, 0, n, k * C3_RUN);  // contig holding the run's first slot
    c_hi[k] = (int32_t)last_le(off, 0, n, B - 1 < off[n] ? B - 1 : off[n]);
    c_ev[k] = (int32_t)last_le(ev_off, 0, n, e_lo[k]);  // contig owning the run's first event (any, if there is none)
    if (weight_count) atomicAdd(&weight_count[run_weight_class(e_hi[k] - e_lo[k])], 1u);
}

// Runs are drawn heaviest first.  A run's cost grows with its number of events (a run inside an organism at 1500x
// holds 60 times the events of one at 25x and takes one warp ~15 times as long); drawn in array order, such runs at
// the end of the array were a tail of several milliseconds with the rest of the device idle.  A counting sort by weight
// class (order inside a class is whatever the atomics give: every result is an order-independent integer sum).
__global__ void cov_run_order_kernel(const int64_t *__restrict__ e_lo, const int64_t *__restrict__ e_hi, int64_t n_runs,
                                     const unsigned int *__restrict__ weight_count, unsigned int *__restrict__ fill,
                                     int32_t *__restrict__ order, const int *__restrict__ gate) {
    __shared__ unsigned int s_first[C4_WEIGHT_CLASSES];
    if (*gate) return;
    if (threadIdx.x == 0) {
        unsigned int acc = 0;
        for (int b = 0; b < C4_WEIGHT_CLASSES; b++) { s_first[b] = acc; acc += weight_count[b]; }
    }
    __syncthreads();
    const int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (k >= n_runs) return;
    const int b = run_weight_class(e_hi[k] - e_lo[k]);
    order[s_first[b] + atomicAdd(&fill[b], 1u)] = (int32_t)k;
}

// Out-of-line pieces of cov_stream_kernel: they run once per contig (or never), keeping them out of the
// streaming loop keeps the loop inside the instruction cache.
__device__ __noinline__ void c3_deep_chunk(const TileArgs &a, int64_t c, const int32_t *chunk, int depth_before,
                                           uint32_t m) {
    int dd[COV_ITEMS];
    int run_d = depth_before;
    for (int i = 0; i < 16; i++) { run_d += chunk[i]; dd[i] = run_d; }
    ct_deep_positions(a, c, dd, m);
}

// The window of contig c is complete (or the run ends): finish it from the warp's histogram when it
// lies inside this run, else merge into the row of the run it started in (last contributor finishes).
__device__ __noinline__ void c3_close_window(const TileArgs &a, int64_t c, bool local, int maxd, int in_run,
                                             uint32_t *s_hist) {
    const int lane = threadIdx.x & 31;
    const int64_t excl = a.tp.excl;
    const int64_t o0g = a.off[c], o1g = a.off[c + 1];
    const uint64_t N = (uint64_t)(o1g - o0g - 2 * excl);
    const uint32_t mtop = (uint32_t)(maxd < COV_BINS - 1 ? maxd : COV_BINS - 1);
    __syncwarp();
    if (local) {
        finish_contig_t(a, c, N, (uint32_t)maxd, [&](uint32_t i) { return s_hist[i]; });
    } else {
        const int64_t row = (o0g + excl) / C3_RUN;
        uint32_t *r = a.rows + row * COV_BINS;
        for (uint32_t i = lane; i <= mtop; i += 32) {
            const uint32_t v = s_hist[i];
            if (v) atomicAdd(&r[i], v);
        }
        __threadfence();
        __syncwarp();
        int last = 0;
        if (lane == 0) {
            atomicMax(&a.meta[row].maxd, (unsigned)maxd);
            __threadfence();
            const unsigned long long old = atomicAdd(&a.meta[row].done, (unsigned long long)in_run);
            last = (old + (unsigned long long)in_run == N);
        }
        last = __shfl_sync(0xffffffffu, last, 0);
        if (last) {
            __threadfence();
            const uint32_t gm = *((volatile unsigned int *)&a.meta[row].maxd);
            finish_contig_t(a, c, N, gm, [&](uint32_t i) { return __ldcg(r + i); });
        }
    }
    __syncwarp();
    for (uint32_t i = lane; i <= mtop; i += 32) s_hist[i] = 0;
    __syncwarp();
}

// contig cursor of the event stream as a register quad: x = contig index, y = first event of the following
// contig (relative to the run's first event, saturated), z = contig start - run_lo (saturated), w = length
__device__ __noinline__ int4 c3_seek(const int64_t *__restrict__ ev_off, const int64_t *__restrict__ off, int ec,
                                     int c_hi_excl, int64_t e_abs, int64_t e_base, int64_t run_lo) {
    const int64_t c = last_le(ev_off, ec, c_hi_excl, e_abs);
    const int64_t o0 = off[c];
    return make_int4((int)c, clamp_rel(ev_off[c + 1] - e_base), clamp_rel(o0 - run_lo), (int)(off[c + 1] - o0));
}

#ifndef MP2_C3_WPC
#define MP2_C3_WPC 1
#endif
constexpr int C3_WPC = MP2_C3_WPC;  // warps per CTA: each works alone on its own slice of the shared memory; more
                                    // than one only spreads the 1 KB the system reserves per CTA over more warps

__global__ void __launch_bounds__(32 * C3_WPC)
cov_stream_kernel(const __grid_constant__ TileArgs a) {
    __shared__ __align__(16) int32_t s_T_all[C3_WPC][C3_WORDS];
    __shared__ uint32_t s_hist_all[C3_WPC][COV_BINS];
    int32_t *s_T = s_T_all[threadIdx.x >> 5];
    uint32_t *s_hist = s_hist_all[threadIdx.x >> 5];
    if (a.check && a.check[0] != 0) {
        if (blockIdx.x == 0 && threadIdx.x == 0) *a.gate_out = 1;
        return;
    }
    if (*((volatile int *)a.gate_out)) return;
    const int lane = threadIdx.x & 31;
    const uint32_t h0 = smem_base_opaque(s_hist), sT0 = smem_base_opaque(s_T);
    const int excl = (int)a.tp.excl;
    const uint32_t span = a.span_dev ? *a.span_dev : a.span_host;  // <= C3_S (cov_bounds_run_kernel gates otherwise)
    for (int i = lane; i < COV_BINS; i += 32) s_hist[i] = 0;
    __syncwarp();

    for (;;) {
        int64_t run = 0;
        if (lane == 0) run = atomicAdd(a.work_counter, 1);
        run = __shfl_sync(0xffffffffu, run, 0);
        if (run >= a.n_tiles) break;
        const int64_t run_lo = run * C3_RUN;
        const int rlen = (int)(run_lo + C3_RUN < a.n_pos ? C3_RUN : a.n_pos - run_lo);
        const int64_t e_base = a.e_lo[run];                 // first event this run looks at
        const int n_ev = (int)(a.e_hi[run] - e_base);        // events of the run (fits 32 bits)
        const uint32_t *__restrict__ ev_s = a.starts + e_base;
        const uint32_t *__restrict__ ev_e = a.ends + e_base;
        int e_cur = 0;                                       // next event to consume, relative to e_base
        const int64_t c_last = a.c_hi[run];
        const int c_hi_excl = (int)(c_last + 1 < a.n ? c_last + 1 : a.n);
        int4 cur = make_int4(0, 0, 0, 0);
        if (n_ev > 0) cur = c3_seek(a.ev_off, a.off, 0, (int)a.n, e_base, e_base, run_lo);
        // events are fetched in groups of 128 (4 coalesced batches of 32), one group ahead of the one being
        // applied: 16 loads in flight per lane cover the HBM latency without relying on occupancy
        uint32_t cs[4], ce[4], ns[4], ne[4];
#define C3_LOAD_GROUP(S, E, BASE)                                               \
    _Pragma("unroll") for (int j_ = 0; j_ < 4; j_++) {                          \
        const int i_ = (BASE) + 32 * j_ + lane;                                 \
        S[j_] = 0; E[j_] = 0;                                                   \
        if (i_ < n_ev) { S[j_] = __ldg(ev_s + i_); E[j_] = __ldg(ev_e + i_); }  \
    }
        C3_LOAD_GROUP(cs, ce, e_cur);
        C3_LOAD_GROUP(ns, ne, e_cur + 128);
        for (int i = lane; i < C3_WORDS / 4; i += 32) reinterpret_cast<int4 *>(s_T)[i] = make_int4(0, 0, 0, 0);
        __syncwarp();

        // window state of the contig being histogrammed
        int64_t c = a.c_lo[run];
        bool have = false;
        int wlo = 0, whi = -1;      // run-relative window of contig c
        int in_run = 0;             // window slots of c accumulated in this run
        int maxd = 0;               // largest depth of c seen in this run (warp-uniform)
        int carry = 0;              // depth just before the current window
        bool viol = false;          // a block longer than the span the event ranges were computed with

        for (int k = 0; k < C3_TPR; k++) {
            const int t_lo = k * C3_W;
            if (t_lo >= rlen) break;
            const int t_hi = t_lo + C3_W < rlen ? t_lo + C3_W : rlen;
            // ---- events whose start lies before t_hi ----
            int open = 0;
            for (;;) {
                bool partial = false;
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    if (!partial) {
                        const int e = e_cur + lane;
                        const bool live = e < n_ev;
                        const uint32_t sv = cs[j], ev = ce[j];
                        int crel = cur.z;
                        uint32_t len = (uint32_t)cur.w;
                        if (live && e >= cur.y) {  // this lane's event belongs to a later contig (rare)
                            const int4 mine = c3_seek(a.ev_off, a.off, cur.x, c_hi_excl, e_base + e, e_base, run_lo);
                            crel = mine.z;
                            len = (uint32_t)mine.w;
                        }
                        const uint32_t en = ev < len ? ev : len;
                        const bool ok = live && sv < len && en > sv;
                        viol |= ok && (en - sv > span);  // a vouched max_span that does not hold (checked per run)
                        const int gs = (int)((uint32_t)crel + sv), ge = (int)((uint32_t)crel + en);  // run-relative
                        // consumed now: start < t_hi (a start beyond its contig is dropped at once, like the general path)
                        const bool before = live && (sv >= len || gs < t_hi);
                        const unsigned nb = ~__ballot_sync(0xffffffffu, before);
                        const int cnt = nb ? __ffs(nb) - 1 : 32;  // leading lanes only (starts are sorted)
                        {   // predicated reductions: no divergent branches in the streaming loop
                            const bool act = ok && lane < cnt;
                            const int rs = gs - t_lo, re = ge - t_lo;
                            smem_inc_if(sT0 + 4u * (uint32_t)c3_phys(rs), act && rs >= 0);
                            open += (act && rs < 0 && re >= 0);  // first window of a run: block open across run_lo
                            smem_dec_if(sT0 + 4u * (uint32_t)c3_phys(re), act && re >= 0 && re < C3_W + C3_S);
                        }
                        e_cur += cnt;
                        if (cnt && e_cur < n_ev && e_cur >= cur.y)  // advance the warp's contig cursor
                            cur = c3_seek(a.ev_off, a.off, cur.x, c_hi_excl, e_base + e_cur, e_base, run_lo);
                        if (cnt < 32) partial = true;
                    }
                }
                if (!partial) {  // steady state: the prefetched group becomes current, fetch the next one
#pragma unroll
                    for (int j = 0; j < 4; j++) { cs[j] = ns[j]; ce[j] = ne[j]; }
                    C3_LOAD_GROUP(ns, ne, e_cur + 128);
                    // and pull the group three ahead into L2 (no register, no scoreboard): lanes 0-3 touch the
                    // four 128-byte lines of its starts, lanes 4-7 those of its ends
                    if (lane < 8) {
                        const int i_ = e_cur + C3_PF + 32 * (lane & 3);
                        if (i_ < n_ev) {
                            const uint32_t *p_ = (lane < 4 ? ev_s : ev_e) + i_;
                            asm volatile("prefetch.global.L2 [%0];" ::"l"(p_));
                        }
                    }
                } else {         // window done: restart at the first unconsumed event; loads overlap the scan below
                    C3_LOAD_GROUP(cs, ce, e_cur);
                    C3_LOAD_GROUP(ns, ne, e_cur + 128);
                    break;
                }
            }
            open = __reduce_add_sync(0xffffffffu, open);
            carry += open;
            __syncwarp();

            // ---- this lane's 64 slots: total, exclusive warp scan ----
            const int myb = (C3_SPL + 4) * lane;
            int tot = 0;
            {
                const int4 *p = reinterpret_cast<const int4 *>(s_T + myb);
#pragma unroll
                for (int j = 0; j < C3_SPL / 4; j++) {
                    const int4 v = p[j];
                    tot += v.x + v.y + v.z + v.w;
                }
            }
            int incl = tot;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += t;
            }
            const int dstart = carry + incl - tot;
            carry += __shfl_sync(0xffffffffu, incl, 31);

            // ---- contig windows intersecting [t_lo, t_hi) ----
            const int q0 = t_lo + C3_SPL * lane;
            for (;;) {
                if (!have) {
                    if (c > c_last) break;
                    const int64_t o0 = a.off[c], o1 = a.off[c + 1];
                    const int r0 = clamp_rel(o0 - run_lo);
                    if (r0 >= t_hi) break;
                    const int r1 = clamp_rel(o1 - run_lo);
                    if (o1 - o0 > 2 * (int64_t)excl) { wlo = r0 + excl; whi = r1 - excl - 1; }
                    else { wlo = 0; whi = -1; }
                    have = true;
                    in_run = 0;
                    maxd = 0;
                }
                if (whi < wlo || whi < 0) { c++; have = false; continue; }  // no window / ended before this run
                if (wlo >= t_hi) break;
                const int lo = wlo > t_lo ? wlo : t_lo, hi = whi + 1 < t_hi ? whi + 1 : t_hi;
                if (hi > lo) {
                    int mxl = 0;
                    if (lo < q0 + C3_SPL && hi > q0) {  // this lane's slots intersect the window
                        int run_d = dstart;
#pragma unroll 1
                        for (int ch = 0; ch < C3_SPL / 16; ch++) {
                            const int q = q0 + 16 * ch;
                            const int d_before = run_d;
                            int d[16];
                            const int4 *p = reinterpret_cast<const int4 *>(s_T + myb + 16 * ch);
#pragma unroll
                            for (int j = 0; j < 4; j++) {
                                const int4 v = p[j];
                                run_d += v.x; d[4 * j] = run_d;
                                run_d += v.y; d[4 * j + 1] = run_d;
                                run_d += v.z; d[4 * j + 2] = run_d;
                                run_d += v.w; d[4 * j + 3] = run_d;
                            }
                            const int l2 = lo > q ? lo : q, h2 = hi < q + 16 ? hi : q + 16;
                            if (h2 > l2) {
                                const uint32_t m = ((1u << (h2 - q)) - 1u) & ~((1u << (l2 - q)) - 1u);
                                int mk = 0;
                                if (m == 0xFFFFu) {
#pragma unroll
                                    for (int i = 0; i < 16; i++) mk = d[i] > mk ? d[i] : mk;
                                    if (mk < COV_BINS - 1) {
#pragma unroll
                                        for (int i = 0; i < 16; i++) smem_inc(h0 + 4u * (uint32_t)d[i]);
                                    } else {
#pragma unroll
                                        for (int i = 0; i < 16; i++)
                                            if (d[i] < COV_BINS - 1) smem_inc(h0 + 4u * (uint32_t)d[i]);
                                    }
                                } else {
#pragma unroll
                                    for (int i = 0; i < 16; i++) {
                                        const uint32_t in = m & (1u << i);
                                        mk = (in && d[i] > mk) ? d[i] : mk;
                                        if (in && d[i] < COV_BINS - 1) smem_inc(h0 + 4u * (uint32_t)d[i]);
                                    }
                                }
                                if (mk >= COV_BINS - 1) c3_deep_chunk(a, c, s_T + myb + 16 * ch, d_before, m);  // rare
                                mxl = mk > mxl ? mk : mxl;
                            }
                        }
                    }
                    mxl = __reduce_max_sync(0xffffffffu, mxl);
                    maxd = mxl > maxd ? mxl : maxd;
                    in_run += hi - lo;
                }
                if (whi >= t_hi) break;  // the window continues in the next 2048 slots
                // ---- the window of contig c ends here ----
                c3_close_window(a, c, wlo >= 0, maxd, in_run, s_hist);
                c++;
                have = false;
            }
            // ---- slide: the overflow region becomes the head of the next window ----
            __syncwarp();
            if (k + 1 < C3_TPR && t_hi < rlen) {
                const int src = c3_phys(C3_W);  // 2176
                int4 keep[(C3_HEAD / 4 + 31) / 32];
#pragma unroll
                for (int j = 0; j < (C3_HEAD / 4 + 31) / 32; j++) {
                    const int i = lane + 32 * j;
                    keep[j] = i < C3_HEAD / 4 ? reinterpret_cast<const int4 *>(s_T + src)[i] : make_int4(0, 0, 0, 0);
                }
                __syncwarp();
                for (int i = lane; i < C3_WORDS / 4; i += 32) reinterpret_cast<int4 *>(s_T)[i] = make_int4(0, 0, 0, 0);
                __syncwarp();
#pragma unroll
                for (int j = 0; j < (C3_HEAD / 4 + 31) / 32; j++) {
                    const int i = lane + 32 * j;
                    if (i < C3_HEAD / 4) reinterpret_cast<int4 *>(s_T)[i] = keep[j];
                }
                __syncwarp();
            }
        }
        // ---- run end: a window still open is merged into the row of the run it started in ----
        if (have && whi >= wlo && in_run > 0) c3_close_window(a, c, false, maxd, in_run, s_hist);
        if (__any_sync(0xffffffffu, viol) && lane == 0) *a.gate_out = 1;  // unreliable ranges: general path redoes the sample
    }
}


// =====================================================================================================
// cov_stream2_kernel: the streaming algorithm again, leaner and order-tolerant.
//
//   * Events reach the warp through a ring in its shared memory filled by asynchronous copies (cp.async =
//     LDGSTS, 4 stages of 64 events per array, one commit group per stage): no registers, no scoreboard and
//     no re-alignment when a window ends in the middle of a batch -- the next batch simply starts at e_cur.
//     (A first version used the bulk-copy engine, cp.async.bulk + mbarrier: UBLKCP takes its operands from
//     uniform registers, and since the compiler cannot prove the per-run addresses warp-uniform it wraps every
//     copy in an elect / R2UR.BROADCAST / retry loop -- about 100 instructions per stage, a third of all
//     instructions of the kernel.  Per-lane 4-byte copies cost 12.)
//   * Bounded disorder: the blocks of a BAM are ordered by READ position, so the second block of a read with
//     a D / N operation starts after the first blocks of the following reads.  Events are consumed by the
//     running maximum of their starts: a batch is taken whole while max(pm, batch max) < t_hi + delta
//     (one REDUX per batch); only the batch that crosses the limit is split by an exact prefix maximum.
//     Since a start is never more than `disorder` (<= delta) behind the running maximum, every event that
//     starts inside the window has been applied when the window is histogrammed; events taken early land
//     in the overflow region (delta + span <= 512 slots) that becomes the head of the next window.
//   * The difference array holds 4 * (+-1), so the running value of the lane-serial walk IS the byte offset
//     of the depth's histogram bin: one IADD and one ATOMS per position.  The overflow test of 16 positions
//     is one OR chain; the highest occupied bin is looked up when the contig closes instead of being
//     tracked per position.
// =====================================================================================================
#ifndef MP2_C4_STAGE
#define MP2_C4_STAGE 64
#endif
#ifndef MP2_C4_NST
#define MP2_C4_NST 4
#endif
constexpr int C4_STAGE = MP2_C4_STAGE;       // events per bulk copy and array (a power of two, >= 64)
constexpr int C4_NST = MP2_C4_NST;           // stages = mbarriers (a power of two)
static_assert((C4_STAGE & (C4_STAGE - 1)) == 0 && (C4_NST & (C4_NST - 1)) == 0 && C4_STAGE >= 64, "ring geometry");
constexpr int C4_RING = C4_STAGE * C4_NST;   // events per array in the ring
constexpr int C4_LIMIT4 = 4 * (COV_BINS - 1);

// The histogram sits at a shared-window address that is a multiple of its own size (2 KB; the window of a CTA
// starts with 1 KB the system reserves): base | offset == base + offset for every offset inside it, which the
// position walk uses to test 16 bins with one OR chain.  (Checked when the kernel starts.)
constexpr int C4_PAD = (2048 - (1024 + (int)sizeof(int32_t) * C3_WORDS) % 2048) % 2048;
struct __align__(1024) C4Smem {
    int32_t T[C3_WORDS];         // 4 * difference array: window + overflow region, 4 pad words per 64 slots
    uint8_t pad[C4_PAD];
    uint32_t hist[COV_BINS];
    uint32_t ring_s[C4_RING];
    uint32_t ring_e[C4_RING];
};
static_assert(sizeof(uint32_t) * COV_BINS == 2048 && (1024 + offsetof(C4Smem, hist)) % 2048 == 0, "histogram alignment");
// Shared-window address of the histogram: the window of a CTA starts with 1 KB the system reserves and C4Smem is the
// kernel's only shared object.  Known at compile time, it is the immediate of the histogram's red.shared, so the
// position walk runs on plain bin OFFSETS (checked against the real address when the kernel starts).
constexpr uint32_t C4_HIST_ADDR = 1024 + (uint32_t)offsetof(C4Smem, hist);
__device__ __forceinline__ void c4_hist_inc(uint32_t byte_off) {
    asm volatile("red.shared.add.u32 [%0+%1], 1;" ::"r"(byte_off), "n"(C4_HIST_ADDR) : "memory");
}

__device__ __forceinline__ uint32_t lds_u32(uint32_t saddr) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(saddr) : "memory");
    return v;
}
__device__ __forceinline__ int4 lds_v4(uint32_t saddr) {
    int4 v;
    asm volatile("ld.shared.v4.s32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(saddr) : "memory");
    return v;
}

// cp.async (LDGSTS): global -> shared without a register in between; 4-byte copies take any alignment
__device__ __forceinline__ void cp_async4(uint32_t dst, const void *src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int V> __device__ __forceinline__ void smem_add_lit(uint32_t saddr) {
    asm volatile("red.shared.add.s32 [%0], %1;" ::"r"(saddr), "n"(V) : "memory");
}
template <int N> __device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void smem_add_if(uint32_t saddr, int v, bool cond) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %2, 0;\n\t@p red.shared.add.s32 [%0], %1;\n\t}" ::"r"(saddr), "r"(v),
                 "r"((uint32_t)cond)
                 : "memory");
}

// contig cursor (see c3_seek) of event e_abs, which belongs to contig c0 or a later one.  Nearly always it is
// the NEXT contig that has events, so the four numbers of contigs c0+1 / c0+2 are fetched at once (one memory
// latency instead of a chain of five dependent loads: a seek happens at every contig boundary of the event
// stream, and the warp does nothing else while it waits); anything else gallops on from there.
__device__ __noinline__ int4 c4_seek(const int64_t *__restrict__ ev_off, const int64_t *__restrict__ off, int c0, int64_t n,
                                     int64_t e_abs, int64_t e_base, int64_t run_lo) {
    int64_t lo = c0;
    if (lo + 2 <= n) {
        const int64_t e1 = __ldg(ev_off + lo + 1), e2 = __ldg(ev_off + lo + 2);
        const int64_t o1 = __ldg(off + lo + 1), o2 = __ldg(off + lo + 2);
        if (e1 <= e_abs && e_abs < e2)
            return make_int4((int)(lo + 1), clamp_rel(e2 - e_base), clamp_rel(o1 - run_lo), (int)(o2 - o1));
        if (e1 <= e_abs) lo += 2;  // e2 <= e_abs as well
    }
    int64_t step = 1;
    while (lo + step < n && ev_off[lo + step] <= e_abs) { lo += step; step <<= 1; }
    const int64_t hi = lo + step < n ? lo + step : n;
    const int64_t c = last_le(ev_off, lo, hi, e_abs);
    const int64_t o0 = off[c];
    return make_int4((int)c, clamp_rel(ev_off[c + 1] - e_base), clamp_rel(o0 - run_lo), (int)(off[c + 1] - o0));
}

// ---- depths outside the warp's 511 bins -----------------------------------------------------------------------
// The shared histogram covers the depths [base, base + 511) of the contig being walked; `base` starts at 0 and
// FOLLOWS the contig's depth (c4_rebase), so an organism at 700x is histogrammed in shared memory like one at 20x.
// Everything else -- positions outside the current range, and the bins of a range that was left -- is counted at
// its ABSOLUTE depth in the contig's row of the deep pool (claimed on first use; CT_DEEP_BINS depths).  A contig's
// histogram is then pool[d] + (d in range ? shared[d - base] : 0).  Round 1 (and the first version of this kernel)
// kept base = 0 and sent every position at depth >= 511 to the pool with a global atomic of its own: exact, but a
// sample with a few very abundant organisms ran 20 % slower than its neighbours (C4 on 8 GPUs).
constexpr int C4_RANGE = COV_BINS - 1;  // bins 0 .. 510 are depths base .. base + 510

// row of the deep pool for contig c (claimed if need be), by ONE lane; -3 = pool exhausted
__device__ __forceinline__ int c4_claim_row(const TileArgs &a, int64_t c) {
    int r = atomicAdd(&a.deep_row_of[c], 0);
    if (r == -1) {
        const int old = atomicCAS(&a.deep_row_of[c], -1, -2);
        if (old == -1) {
            const int k2 = atomicAdd(a.deep_rows_used, 1);
            r = k2 < a.deep_rows ? k2 : -3;
            atomicExch(&a.deep_row_of[c], r);
        } else {
            r = old;
        }
    }
    while (r == -2) r = atomicAdd(&a.deep_row_of[c], 0);
    return r;
}

// 16 positions that are not all inside the contig window, or not all inside the histogram's depth range: exact,
// one by one.  chunk = 4 * differences, d4_before = 4 * (absolute) depth just before the chunk, m = positions
// inside the window.  Called by any subset of the lanes.  Returns the largest depth among the positions that went to
// the POOL (-1: none did; the largest depth inside the shared bins is read off the bins when the contig closes).
__device__ __noinline__ int c4_slow_chunk(const TileArgs &a, int64_t c, const int32_t *chunk, int d4_before, uint32_t m,
                                          int base, uint32_t *hist) {
    const int lane = threadIdx.x & 31;
    int dd[COV_ITEMS];
    int run = d4_before, mx = -1;
    uint32_t out = 0;  // positions whose depth is outside [base, base + 511)
#pragma unroll
    for (int i = 0; i < 16; i++) {
        run += chunk[i];
        int d = run >> 2;
        d = d < 0 ? 0 : d;
        dd[i] = d;
        if (m & (1u << i)) {
            if ((unsigned)(d - base) < (unsigned)C4_RANGE) {
                atomicAdd(&hist[d - base], 1u);
            } else {
                out |= 1u << i;
                mx = d > mx ? d : mx;
            }
        }
    }
    if (out) {
        const unsigned grp = __activemask();
        const int leader = __ffs(grp) - 1;
        int r = 0;
        if (lane == leader) r = c4_claim_row(a, c);
        r = __shfl_sync(grp, r, leader);
        if (r >= 0) {
            uint32_t *dr = a.deep_pool + (size_t)r * CT_DEEP_BINS;
            for (int k = 0; k < COV_ITEMS; k++)
                if ((out & (1u << k)) && dd[k] < CT_DEEP_BINS) atomicAdd(&dr[dd[k]], 1u);  // deeper: the finish gates
            __threadfence();
        }
    }
    return mx;
}

// Empties the shared histogram (range [base, base + 511)) into the pool at the bins' absolute depths, so that the
// caller can give it another range.  Whole warp.  Returns the largest depth that has left shared memory so far
// (warp-uniform; arguments and result by value: references would live on the stack).
__device__ __noinline__ int c4_rebase(const TileArgs &a, int64_t c, uint32_t *hist, int base, int deep_max) {
    const int lane = threadIdx.x & 31;
    __syncwarp();
    uint32_t v[COV_BINS / 32];
    int top = -1;
#pragma unroll
    for (int j = 0; j < COV_BINS / 32; j++) {
        v[j] = hist[lane + 32 * j];
        if (v[j]) top = lane + 32 * j;
    }
    top = __reduce_max_sync(0xffffffffu, top);
    deep_max = __reduce_max_sync(0xffffffffu, deep_max);
    if (top >= 0) {
        int r = 0;
        if (lane == 0) r = c4_claim_row(a, c);
        r = __shfl_sync(0xffffffffu, r, 0);
        if (r >= 0) {
            uint32_t *dr = a.deep_pool + (size_t)r * CT_DEEP_BINS;
#pragma unroll
            for (int j = 0; j < COV_BINS / 32; j++) {
                const int d = base + lane + 32 * j;
                if (v[j] && d < CT_DEEP_BINS) atomicAdd(&dr[d], v[j]);
            }
            __threadfence();
        }
        deep_max = base + top > deep_max ? base + top : deep_max;
#pragma unroll
        for (int j = 0; j < COV_BINS / 32; j++)
            if (v[j]) hist[lane + 32 * j] = 0;
    }
    __syncwarp();
    return deep_max;
}

// Trimmed mean of contig c from `local(d)` (this warp's shared bins or the contig's global row) + its pool row.
// use_pool: somebody counted positions of this contig in the pool (otherwise its row is not even looked up).
template <class Local>
__device__ __forceinline__ void c4_finish(const TileArgs &a, int64_t c, uint64_t N, uint32_t maxd, bool use_pool, Local local) {
    const int lane = threadIdx.x & 31;
    const int r = use_pool ? *((volatile int32_t *)&a.deep_row_of[c]) : -1;
    if (maxd >= (uint32_t)CT_DEEP_BINS || r == -3 || r == -2) {  // beyond the pool's range / pool exhausted
        if (lane == 0) {
            a.deep_list[atomicAdd(a.deep_count, 1)] = (int32_t)c;
            *a.gate_out = 1;  // the exact general path redoes this sample
        }
        return;
    }
    const uint32_t *deep = r >= 0 ? a.deep_pool + (size_t)r * CT_DEEP_BINS : nullptr;
    auto cnt = [&](uint32_t d) { return local(d) + (deep ? __ldcg(deep + d) : 0u); };
    uint64_t mn, mx;
    trim_indices(N, a.tp, mn, mx);
    const uint64_t covered = N - (uint64_t)cnt(0);
    unsigned long long acc = 0, total = 0;
    if (covered) trim_walk_warp(cnt, 0, maxd + 1, mn, mx, acc, total);
    if (lane == 0) write_result(c, N, covered, mn, mx, total, maxd, 0, a.cov, a.cov_stride, a.stats);
}

// The window of contig c is complete (or the run ends).  Inside this run: finish from the warp's bins (+ pool).
// Otherwise: bring the bins back to base 0 semantics (what is not in [0, 511) goes to the pool), merge them into the
// row of the run the window started in; the last contributor finishes from that row (+ pool).
__device__ __noinline__ void c4_close_window(const TileArgs &a, int64_t c, bool local, int deep_max, int in_run, int base,
                                             uint32_t *s_hist) {
    const int lane = threadIdx.x & 31;
    const int64_t excl = a.tp.excl;
    const int64_t o0g = a.off[c], o1g = a.off[c + 1];
    const uint64_t N = (uint64_t)(o1g - o0g - 2 * excl);
    __syncwarp();
    if (!local && base != 0) {
        deep_max = c4_rebase(a, c, s_hist, base, deep_max);
        base = 0;
    }
    int top = -1;
#pragma unroll
    for (int j = 0; j < COV_BINS / 32; j++)  // lane + 32 j: conflict-free
        if (s_hist[lane + 32 * j]) top = lane + 32 * j;
    top = __reduce_max_sync(0xffffffffu, top);
    deep_max = __reduce_max_sync(0xffffffffu, deep_max);
    const bool pooled = deep_max >= 0;  // this warp put positions of c into the pool (deep_max starts at -1)
    int maxd = top >= 0 && base + top > deep_max ? base + top : deep_max;
    maxd = maxd < 0 ? 0 : maxd;
    if (local) {
        const int b0 = base;
        c4_finish(a, c, N, (uint32_t)maxd, pooled, [&](uint32_t d) {
            return (d - (uint32_t)b0) < (uint32_t)C4_RANGE ? s_hist[d - (uint32_t)b0] : 0u;
        });
    } else {
        const int64_t row = (o0g + excl) / C3_RUN;
        uint32_t *r = a.rows + row * COV_BINS;
        for (int i = lane; i <= top; i += 32) {
            const uint32_t v = s_hist[i];
            if (v) atomicAdd(&r[i], v);
        }
        __threadfence();
        __syncwarp();
        int last = 0;
        if (lane == 0) {
            atomicMax(&a.meta[row].maxd, (unsigned)maxd);
            if (pooled) atomicOr(&a.meta[row].pad, 1u);
            __threadfence();
            const unsigned long long old = atomicAdd(&a.meta[row].done, (unsigned long long)in_run);
            last = (old + (unsigned long long)in_run == N);
        }
        last = __shfl_sync(0xffffffffu, last, 0);
        if (last) {
            __threadfence();
            const uint32_t gm = *((volatile unsigned int *)&a.meta[row].maxd);
            const bool any_pooled = *((volatile unsigned int *)&a.meta[row].pad) != 0;
            c4_finish(a, c, N, gm, any_pooled, [&](uint32_t d) { return d < (uint32_t)C4_RANGE ? __ldcg(r + d) : 0u; });
        }
    }
    __syncwarp();
    for (int i = lane; i <= top; i += 32) s_hist[i] = 0;
    __syncwarp();
}

__global__ void __launch_bounds__(32)
cov_stream2_kernel(const __grid_constant__ TileArgs a) {
    __shared__ C4Smem sm;
    if (*((volatile int *)a.gate_out)) return;
    const int lane = threadIdx.x;
    const uint32_t sT0 = smem_base_opaque(sm.T);
    const uint32_t rS0 = smem_base_opaque(sm.ring_s), rE0 = smem_base_opaque(sm.ring_e);
    const uint32_t h0 = smem_base_opaque(sm.hist);
    if (h0 != C4_HIST_ADDR) {  // the shared window is not laid out as assumed: leave the sample to the general path
        if (lane == 0) *a.gate_out = 1;
        return;
    }
    const int excl = (int)a.tp.excl;
    const uint32_t span = a.span_dev ? *a.span_dev : a.span_host;
    // events are consumed up to `delta` slots ahead of the window; delta >= disorder and delta + span <= C3_S
    // (cov_bounds_run_kernel sends anything else to the general path)
    const int delta = C3_S - (int)(span < (uint32_t)C3_S ? span : (uint32_t)C3_S);
    for (int i = lane; i < COV_BINS; i += 32) sm.hist[i] = 0;
    __syncwarp();

    for (;;) {
        int64_t run = 0;
        if (lane == 0) run = atomicAdd(a.work_counter, 1);
        run = __shfl_sync(0xffffffffu, run, 0);
        if (run >= a.n_tiles) break;
        if (a.order) run = a.order[run];
        const int64_t run_lo = run * C3_RUN;
        const int rlen = (int)(run_lo + C3_RUN < a.n_pos ? C3_RUN : a.n_pos - run_lo);
        const int64_t e_base = a.e_lo[run];
        const int64_t n_ev64 = a.e_hi[run] - e_base;
        const int n_ev = n_ev64 > 0 ? (int)n_ev64 : 0;
        // ---- staging: stage s = events [ST s, ST s + ST) of the run, in ring slot s % NST; one commit group per
        // stage, and always floor(e_cur / ST) + NST groups committed (empty ones beyond the run's last event), so
        // that "all but the NST - 2 most recent groups" covers every stage the next 64 events can be in
        constexpr int ST = C4_STAGE, RM = C4_RING - 1;
        const uint32_t *__restrict__ ev_s = a.starts + e_base;
        const uint32_t *__restrict__ ev_e = a.ends + e_base;
        int issued = 0;   // stages (= groups) committed
        int avail = 0;    // events [0, avail) of the run are in the ring and visible to every lane
        int next_issue = 0;  // the next stage may be issued once e_cur reaches this event: ST * (issued - NST + 1)
        // this lane's next copy: event ST * issued + lane, ring byte offset 4 * (that & RM)
        const uint32_t *p_s = ev_s + lane, *p_e = ev_e + lane;
        uint32_t off = 4u * (uint32_t)lane;
        int e_next = lane;
        auto issue = [&]() {
#pragma unroll
            for (int h = 0; h < ST / 32; h++) {
                if (e_next + 32 * h < n_ev) {
                    cp_async4(rS0 + off + 128u * h, p_s + 32 * h);
                    cp_async4(rE0 + off + 128u * h, p_e + 32 * h);
                }
            }
            cp_async_commit();
            p_s += ST;
            p_e += ST;
            e_next += ST;
            off = (off + 4u * ST) & (uint32_t)(4 * C4_RING - 1);
            issued++;
            next_issue += ST;
        };
        cp_async_wait<0>();
        __syncwarp();  // the previous run's copies have landed and its reads of the ring are done
        for (int s_ = 0; s_ < C4_NST; s_++) issue();
        next_issue = ST;  // stage NST goes into the slot of stage 0: after the first ST events

        int e_cur = 0;  // next event to consume, relative to e_base
        int4 cur = make_int4(0, 0, 0, 0);
        if (n_ev > 0) cur = c4_seek(a.ev_off, a.off, a.c_ev[run], a.n, e_base, e_base, run_lo);
        for (int i = lane; i < C3_WORDS / 4; i += 32) reinterpret_cast<int4 *>(sm.T)[i] = make_int4(0, 0, 0, 0);
        __syncwarp();

        int64_t c = a.c_lo[run];
        const int64_t c_last = a.c_hi[run];
        bool have = false;
        int wlo = 0, whi = -1;      // run-relative window of contig c
        int in_run = 0;             // window slots of c accumulated in this run
        int deep_max = -1;          // largest depth that was counted outside the shared bins for c in this run (per lane; -1: none)
        int base = 0;               // the shared bins are the depths [base, base + 511) of contig c (warp-uniform)
        int carry4 = 0;             // 4 * depth just before the current window
        int pm = INT_MIN;           // running maximum of the starts consumed so far (run-relative)
        // the vouched bounds are checked on three running extremes (per lane) instead of per event:
        int v_re = 0;               // largest window-relative end applied      (must stay < C3_W + C3_S)
        int v_rs = 0;               // smallest window-relative start applied   (must stay >= 0 after the first window)
        uint32_t v_blk = 0;         // longest block                             (must stay <= span)
        bool viol = false;

        for (int k = 0; k < C3_TPR; k++) {
            const int t_lo = k * C3_W;
            if (t_lo >= rlen) break;
            const int t_hi = t_lo + C3_W < rlen ? t_lo + C3_W : rlen;
            const int lim = t_hi + delta;
            int open = 0;
            while (e_cur < n_ev) {
                // ---- staging: refill the slots behind e_cur, wait for the stages the next 64 events are in ----
                if (e_cur >= next_issue) {
                    __syncwarp();  // every lane is done reading the slots that are about to be overwritten
                    do issue(); while (e_cur >= next_issue);
                }
                if ((e_cur + 64 < n_ev ? e_cur + 64 : n_ev) > avail) {
                    cp_async_wait<C4_NST - 2>();  // this lane's copies of all but the NST - 2 newest stages
                    __syncwarp();                 // ... and everybody else's
                    avail = ST * (issued - (C4_NST - 2));
                }
                const int i0 = (e_cur + lane) & RM, j0 = i0;
                const int room = (n_ev < cur.y ? n_ev : cur.y) - e_cur;  // events ahead in this contig
                if (k > 0 && room >= 64) {
                    // ---- 64 events of one contig: both batches are taken whole if none is beyond the limit ----
                    const uint32_t len = (uint32_t)cur.w;
                    const uint32_t sv0 = lds_u32(rS0 + 4u * i0), sv1 = lds_u32(rS0 + 4u * ((i0 + 32) & RM));
                    const uint32_t ev0 = lds_u32(rE0 + 4u * j0), ev1 = lds_u32(rE0 + 4u * ((j0 + 32) & RM));
                    const uint32_t en0 = ev0 < len ? ev0 : len, en1 = ev1 < len ? ev1 : len;
                    const bool ok0 = sv0 < en0, ok1 = sv1 < en1;  // start inside the contig, block not empty
                    const int g0 = cur.z + (int)sv0, g1 = cur.z + (int)sv1;
                    const int key0 = ok0 ? g0 : INT_MIN, key1 = ok1 ? g1 : INT_MIN;
                    int bm = __reduce_max_sync(0xffffffffu, key0 > key1 ? key0 : key1);
                    bm = bm > pm ? bm : pm;
                    if (bm < lim) {
                        // a dropped event becomes (+4, -4) on slot 0 of the window: no predicate for it
                        const int rs0 = ok0 ? g0 - t_lo : 0, rs1 = ok1 ? g1 - t_lo : 0;
                        const uint32_t b0 = ok0 ? en0 - sv0 : 0u, b1 = ok1 ? en1 - sv1 : 0u;
                        const int re0 = rs0 + (int)b0, re1 = rs1 + (int)b1;
                        const bool in0 = rs0 >= 0 && re0 < C3_W + C3_S, in1 = rs1 >= 0 && re1 < C3_W + C3_S;
                        if (in0) {
                            smem_add_lit<4>(sT0 + 4u * (uint32_t)c3_phys(rs0));
                            smem_add_lit<-4>(sT0 + 4u * (uint32_t)c3_phys(re0));
                        }
                        if (in1) {
                            smem_add_lit<4>(sT0 + 4u * (uint32_t)c3_phys(rs1));
                            smem_add_lit<-4>(sT0 + 4u * (uint32_t)c3_phys(re1));
                        }
                        const int mre = re0 > re1 ? re0 : re1, mrs = rs0 < rs1 ? rs0 : rs1;
                        const uint32_t mb = b0 > b1 ? b0 : b1;
                        v_re = mre > v_re ? mre : v_re;
                        v_rs = mrs < v_rs ? mrs : v_rs;
                        v_blk = mb > v_blk ? mb : v_blk;
                        e_cur += 64;
                        pm = bm;
                        continue;
                    }
                }
                // ---- one batch, any shape: contig boundary inside, end of the range, the window's last batch ----
                const int e = e_cur + lane;
                const bool live = e < n_ev;
                const uint32_t sv = lds_u32(rS0 + 4u * i0);
                const uint32_t ev = lds_u32(rE0 + 4u * j0);
                int crel = cur.z;
                uint32_t len = (uint32_t)cur.w;
                if (live && e >= cur.y) {  // this lane's event belongs to a later contig (rare)
                    const int4 mine = c4_seek(a.ev_off, a.off, cur.x, a.n, e_base + e, e_base, run_lo);
                    crel = mine.z;
                    len = (uint32_t)mine.w;
                }
                const uint32_t en = ev < len ? ev : len;
                const bool ok = live && sv < en;
                const int gs = (int)((uint32_t)crel + sv);  // run-relative start
                const int key = ok ? gs : INT_MIN;
                int bm = __reduce_max_sync(0xffffffffu, key);
                bm = bm > pm ? bm : pm;
                int cnt = 32, pm_new = bm;
                if (bm >= lim || room < 32) {  // the batch crosses the limit: exact prefix maximum
                    int p = key;
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) {
                        const int t = __shfl_up_sync(0xffffffffu, p, o);
                        if (lane >= o && t > p) p = t;
                    }
                    p = p > pm ? p : pm;
                    const unsigned nb = ~__ballot_sync(0xffffffffu, live && p < lim);
                    cnt = nb ? __ffs(nb) - 1 : 32;
                    pm_new = __shfl_sync(0xffffffffu, p, cnt ? cnt - 1 : 0);
                    if (!cnt) pm_new = pm;
                }
                const bool act = ok && lane < cnt;
                const int rs = gs - t_lo, re = rs + (int)(en - sv);
                const bool in = act && rs >= 0 && re < C3_W + C3_S;
                viol |= act && (en - sv > span);  // a vouched max_span that does not hold
                if (k == 0) {  // first window of the run: blocks that are open across run_lo
                    const bool tail = act && rs < 0 && re >= 0;
                    open += tail;
                    smem_add_if(sT0 + 4u * (uint32_t)c3_phys(re), -4, tail && re < C3_W + C3_S);
                    viol |= (tail && re >= C3_W + C3_S) || (act && rs >= 0 && !in);
                } else {
                    viol |= act && !in;  // a start before the window was missed earlier; an end beyond the overflow region
                }
                smem_add_if(sT0 + 4u * (uint32_t)c3_phys(rs), 4, in);
                smem_add_if(sT0 + 4u * (uint32_t)c3_phys(re), -4, in);
                e_cur += cnt;
                pm = pm_new;
                if (cnt && e_cur < n_ev && e_cur >= cur.y)  // advance the warp's contig cursor
                    cur = c4_seek(a.ev_off, a.off, cur.x, a.n, e_base + e_cur, e_base, run_lo);
                if (cnt < 32) break;
            }
            open = __reduce_add_sync(0xffffffffu, open);
            carry4 += 4 * open;
            __syncwarp();

            // ---- this lane's 64 slots: total, exclusive warp scan ----
            const int myb = (C3_SPL + 4) * lane;
            const uint32_t my_sa = sT0 + 4u * (uint32_t)myb;  // this lane's 64 slots
            int tot = 0;
#pragma unroll
            for (int j = 0; j < C3_SPL / 4; j++) {
                const int4 v = lds_v4(my_sa + 16u * j);
                tot += v.x + v.y + v.z + v.w;
            }
            int incl = tot;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += t;
            }
            const int dstart4 = carry4 + incl - tot;
            carry4 += __shfl_sync(0xffffffffu, incl, 31);

            // ---- contig windows intersecting [t_lo, t_hi) ----
            const int q0 = t_lo + C3_SPL * lane;
            for (;;) {
                if (!have) {
                    if (c > c_last) break;
                    const int64_t o0 = a.off[c], o1 = a.off[c + 1];
                    const int r0 = clamp_rel(o0 - run_lo);
                    if (r0 >= t_hi) break;
                    const int r1 = clamp_rel(o1 - run_lo);
                    if (o1 - o0 > 2 * (int64_t)excl) { wlo = r0 + excl; whi = r1 - excl - 1; }
                    else { wlo = 0; whi = -1; }
                    have = true;
                    in_run = 0;
                    deep_max = -1;
                    base = 0;
                }
                if (whi < wlo || whi < 0) { c++; have = false; continue; }  // no window / ended before this run
                if (wlo >= t_hi) break;
                const int lo = wlo > t_lo ? wlo : t_lo, hi = whi + 1 < t_hi ? whi + 1 : t_hi;
                if (hi > lo) {
                    const bool mine = lo < q0 + C3_SPL && hi > q0;  // this lane's slots intersect the window
                    // keep the bins' range around the depth, judged by the depth at the start of every lane's slots: one
                    // vote in the common case (nothing near the edges of the range)
                    if (__any_sync(0xffffffffu, mine && (dstart4 >= 4 * (base + 384) || (base > 0 && dstart4 < 4 * (base + 32))))) {
                        const int ds = dstart4 >> 2;
                        const int dmin = __reduce_min_sync(0xffffffffu, mine ? ds : INT_MAX);
                        int nb = dmin - 128;
                        nb = nb < 0 ? 0 : nb;
                        if (nb - base >= 64 || base - nb >= 64) {
                            deep_max = c4_rebase(a, c, sm.hist, base, deep_max);
                            base = nb;
                        }
                    }
                    if (mine) {
                        uint32_t x = (uint32_t)(dstart4 - 4 * base);  // byte offset of the current depth's bin
#pragma unroll 1
                        for (int ch = 0; ch < C3_SPL / 16; ch++) {
                            const int q = q0 + 16 * ch;
                            const uint32_t x_before = x;
                            uint32_t d[16];
#pragma unroll
                            for (int j = 0; j < 4; j++) {
                                const int4 v = lds_v4(my_sa + 64u * ch + 16u * j);
                                x += (uint32_t)v.x; d[4 * j] = x;
                                x += (uint32_t)v.y; d[4 * j + 1] = x;
                                x += (uint32_t)v.z; d[4 * j + 2] = x;
                                x += (uint32_t)v.w; d[4 * j + 3] = x;
                            }
                            uint32_t orr = 0;
#pragma unroll
                            for (int i = 0; i < 16; i++) orr |= d[i];
                            const int l2 = lo > q ? lo : q, h2 = hi < q + 16 ? hi : q + 16;
                            // offsets below the range wrap to huge unsigned values: one OR tests all 16 both ways
                            if (h2 - l2 == 16 && orr < (uint32_t)C4_LIMIT4) {
#pragma unroll
                                for (int i = 0; i < 16; i++) c4_hist_inc(d[i]);
                            } else if (h2 > l2) {
                                const uint32_t m = ((1u << (h2 - q)) - 1u) & ~((1u << (l2 - q)) - 1u);
                                const int mx = c4_slow_chunk(a, c, sm.T + myb + 16 * ch, (int)x_before + 4 * base, m, base, sm.hist);
                                deep_max = mx > deep_max ? mx : deep_max;  // what went to the pool (-1: nothing)
                            }
                        }
                    }
                    in_run += hi - lo;
                }
                if (whi >= t_hi) break;  // the window continues in the next 2048 slots
                c4_close_window(a, c, wlo >= 0, deep_max, in_run, base, sm.hist);
                base = 0;
                c++;
                have = false;
            }
            // ---- slide: the overflow region becomes the head of the next window ----
            __syncwarp();
            if (k + 1 < C3_TPR && t_hi < rlen) {
                const int src = c3_phys(C3_W);
                int4 keep[(C3_HEAD / 4 + 31) / 32];
#pragma unroll
                for (int j = 0; j < (C3_HEAD / 4 + 31) / 32; j++) {
                    const int i = lane + 32 * j;
                    keep[j] = i < C3_HEAD / 4 ? reinterpret_cast<const int4 *>(sm.T + src)[i] : make_int4(0, 0, 0, 0);
                }
                __syncwarp();
                for (int i = lane; i < C3_WORDS / 4; i += 32) reinterpret_cast<int4 *>(sm.T)[i] = make_int4(0, 0, 0, 0);
                __syncwarp();
#pragma unroll
                for (int j = 0; j < (C3_HEAD / 4 + 31) / 32; j++) {
                    const int i = lane + 32 * j;
                    if (i < C3_HEAD / 4) reinterpret_cast<int4 *>(sm.T)[i] = keep[j];
                }
                __syncwarp();
            }
        }
        // ---- run end ----
        if (have && whi >= wlo && in_run > 0) c4_close_window(a, c, false, deep_max, in_run, base, sm.hist);
        viol |= v_re >= C3_W + C3_S || v_rs < 0 || v_blk > span;
        if (__any_sync(0xffffffffu, viol) && lane == 0) *a.gate_out = 1;  // unreliable ranges: the general path redoes the sample

    }
}

}  // namespace mp2

using namespace mp2;

// cudaFuncSetAttribute is per device: done once for every device this process uses
static int cov_tile_attr_once() {
    static std::mutex mu;
    static bool done[64] = {false};
    int dev = 0;
    MP2_CUDA(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lk(mu);
    if (dev >= 0 && dev < 64 && done[dev]) return MP2_OK;
    MP2_CUDA(cudaFuncSetAttribute(cov_tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)CT_SMEM));
    if (dev >= 0 && dev < 64) done[dev] = true;
    return MP2_OK;
}

// MP2_COV_PATH=general forces the difference array through HBM (development / A-B switch).
static bool cov_force_general() {
    static int v = -1;
    if (v < 0) {
        const char *e = getenv("MP2_COV_PATH");
        v = (e && e[0] == 'g') ? 1 : 0;
    }
    return v == 1;
}

// Which fast-path kernel runs: cov_stream2_kernel unless MP2_COV_STREAM=1 (the first streaming kernel) or
// MP2_COV_TILE=1 (block per tile) asks for an older one (development / A-B switches; both need sorted starts).
static int cov_fast_version() {
    static int v = 0;
    if (!v) {
        const char *t = getenv("MP2_COV_TILE"), *o = getenv("MP2_COV_STREAM");
        v = (t && t[0] == '1') ? 1 : (o && o[0] == '1') ? 3 : 4;
    }
    return v;
}

static int cov_general_path(const uint32_t *starts, const uint32_t *ends, const int64_t *ev_off, const int64_t *off,
                            int64_t n_contigs, int64_t n_events, int64_t n_pos, const int32_t *ev_first,
                            const TrimParams &tp, int *ctrs, const int *gate, float *d_cov, int64_t cov_stride,
                            mp2_cov_stats *d_stats, cudaStream_t stream);

// Rows of the deep pool for a call with n_contigs contigs.  An organism sequenced beyond the shared histogram's
// range brings ALL its contigs (hundreds) at once, and how many such organisms a sample holds grows with the shard:
// a C4 shard of 2 500 MAGs (750 k contigs, 4 GPUs) overflowed 2 048 rows in one sample out of ten, which sent those
// samples down the general path (one rank at 464 ms per step, the best one at 303, against 226 without fallbacks).
// One row per 64 contigs, in steps of 1024 rows, 2 048 .. 65 536 (16 KB per row, cleared per sample: 84 MB at C3).
static int deep_pool_rows(int64_t n_contigs) {
    int64_t r = ((n_contigs / 64 + 1023) / 1024) * 1024;
    if (r < CT_DEEP_ROWS) r = CT_DEEP_ROWS;
    if (r > 65536) r = 65536;
    return (int)r;
}

// MP2_COV_RUN_ORDER=0: cov_stream2_kernel draws its runs in array order (development / A-B switch)
static bool cov_run_order_off() {
    static int v = -1;
    if (v < 0) { const char *e = getenv("MP2_COV_RUN_ORDER"); v = (e && e[0] == '0') ? 1 : 0; }
    return v == 1;
}

extern "C" int mp2_cov_events_device(const void *d_starts, const void *d_ends, const void *d_ev_off,
                                     const void *d_contig_off, int64_t n_contigs, int64_t n_events,
                                     int64_t n_positions, uint32_t max_span, uint32_t max_disorder, uint32_t end_excl,
                                     float trim_lower, float trim_upper, void *d_cov, int64_t cov_stride,
                                     void *d_stats, void *stream_) {
    if (n_contigs < 0 || n_events < 0 || n_positions < 0) return fail(MP2_EINVAL, "negative size");
    if (n_contigs > INT32_MAX) return fail(MP2_EINVAL, "too many contigs in one call");
    if (!d_contig_off || !d_ev_off || !d_cov || (n_events > 0 && (!d_starts || !d_ends)))
        return fail(MP2_EINVAL, "NULL argument");
    if ((reinterpret_cast<uintptr_t>(d_starts) & 3) || (reinterpret_cast<uintptr_t>(d_ends) & 3))
        return fail(MP2_EINVAL, "d_starts and d_ends must be 4-byte aligned");
    if (!(trim_lower >= 0.0f) || !(trim_upper >= 0.0f) || !(trim_lower + trim_upper < 1.0f))
        return fail(MP2_EINVAL, "trim_lower and trim_upper must be >= 0 and sum to < 1");
    if (cov_stride < 1) return fail(MP2_EINVAL, "cov_stride must be >= 1");
    MP2_TRY(require_device(-1));
    cudaStream_t stream = (cudaStream_t)stream_;
    if (n_contigs == 0) return MP2_OK;

    const int64_t n_pos = n_positions + 1;
    const int64_t n_ev_tiles = (n_events + COV_EV_TILE - 1) / COV_EV_TILE;
    const int64_t *off = (const int64_t *)d_contig_off;
    const int64_t *ev_off = (const int64_t *)d_ev_off;
    const uint32_t *starts = (const uint32_t *)d_starts, *ends = (const uint32_t *)d_ends;
    TrimParams tp;
    tp.lo = trim_lower;
    tp.hi = 1.0f - trim_upper;  // src/coverage.rs:99
    tp.excl = end_excl;

    // ctrs: [0] deep count (tile path) [1] tile work counter [2] gate [3] deep count (general) [4] general
    // work counter [5] deep next [6] deep rows used; flags: [0] starts go back [1] longest block [2] disorder
    Temp ctrs, flags, ev_first;
    MP2_TRY(ctrs.alloc(sizeof(int) * 8, stream));
    MP2_TRY(flags.alloc(sizeof(unsigned int) * 4, stream));
    MP2_TRY(ev_first.alloc(sizeof(int32_t) * (n_ev_tiles + 1), stream));
    MP2_CUDA(cudaMemsetAsync(ctrs.p, 0, sizeof(int) * 8, stream));
    MP2_CUDA(cudaMemsetAsync(flags.p, 0, sizeof(unsigned int) * 4, stream));
    int *gate = ctrs.as<int>() + 2;
    {
        Launch L("cov_init_kernel", stream);
        cov_init_kernel<<<(unsigned)((n_contigs + 255) / 256), 256, 0, stream>>>(
            off, n_contigs, end_excl, (float *)d_cov, cov_stride, (mp2_cov_stats *)d_stats, gate);
    }
    MP2_CUDA(cudaGetLastError());
    if (n_events > 0) {  // one entry beyond the last tile: the contigs of tile t are first[t] .. first[t + 1]
        Launch L("cov_map_kernel", stream);
        cov_map_kernel<<<(unsigned)((n_ev_tiles + 1 + 255) / 256), 256, 0, stream>>>(ev_off, n_contigs, COV_EV_TILE,
                                                                                   n_ev_tiles + 1, ev_first.as<int32_t>());
    }
    MP2_CUDA(cudaGetLastError());

    const int fast_ver = cov_fast_version();
    const bool need_sorted = fast_ver != 4;
    const bool measure = max_span == 0;
    // a vouched sample the chosen kernel cannot take goes straight to the general path
    bool general_only = cov_force_general() ||
                        (!measure && ((uint64_t)max_span + max_disorder > (uint64_t)C3_S || (need_sorted && max_disorder)));
    // ---------------- fast path: difference array in shared memory ----------------
    if (!general_only) {
        Temp e_lo, e_hi, c_lo, c_hi, c_ev, t_rows, t_meta, t_deep, t_drow, t_dpool, k_max, t_order, t_wcls;
        const int64_t unit = fast_ver == 1 ? CT_TILE : C3_RUN;
        const int64_t n_units = (n_pos + unit - 1) / unit;
        if (n_units > INT32_MAX / 2) return fail(MP2_EINVAL, "too many positions in one call");
        MP2_TRY(e_lo.alloc(sizeof(int64_t) * n_units, stream));
        MP2_TRY(e_hi.alloc(sizeof(int64_t) * n_units, stream));
        MP2_TRY(c_lo.alloc(sizeof(int32_t) * n_units, stream));
        MP2_TRY(c_hi.alloc(sizeof(int32_t) * n_units, stream));
        MP2_TRY(c_ev.alloc(sizeof(int32_t) * n_units, stream));
        MP2_TRY(t_rows.alloc(sizeof(uint32_t) * COV_BINS * (size_t)n_units, stream));
        MP2_TRY(t_meta.alloc(sizeof(RowMeta) * n_units, stream));
        MP2_TRY(t_deep.alloc(sizeof(int32_t) * n_contigs, stream));
        MP2_TRY(t_drow.alloc(sizeof(int32_t) * n_contigs, stream));
        const int deep_rows = deep_pool_rows(n_contigs);
        MP2_TRY(t_dpool.alloc(sizeof(uint32_t) * (size_t)deep_rows * CT_DEEP_BINS, stream));
        MP2_CUDA(cudaMemsetAsync(t_rows.p, 0, sizeof(uint32_t) * COV_BINS * (size_t)n_units, stream));
        MP2_CUDA(cudaMemsetAsync(t_meta.p, 0, sizeof(RowMeta) * n_units, stream));
        MP2_CUDA(cudaMemsetAsync(t_drow.p, 0xFF, sizeof(int32_t) * n_contigs, stream));
        MP2_CUDA(cudaMemsetAsync(t_dpool.p, 0, sizeof(uint32_t) * (size_t)deep_rows * CT_DEEP_BINS, stream));
        const bool ordered = fast_ver == 4 && !cov_run_order_off();
        if (ordered) {  // [0, 64) runs per weight class, [64, 128) fill counters of the counting sort
            MP2_TRY(t_order.alloc(sizeof(int32_t) * n_units, stream));
            MP2_TRY(t_wcls.alloc(sizeof(unsigned int) * 2 * C4_WEIGHT_CLASSES, stream));
            MP2_CUDA(cudaMemsetAsync(t_wcls.p, 0, sizeof(unsigned int) * 2 * C4_WEIGHT_CLASSES, stream));
        }
        if (n_events > 0 && measure) {  // longest block and disorder of the starts, exactly (one pass over the events)
            MP2_TRY(k_max.alloc(sizeof(OrderTile) * n_ev_tiles, stream));
            {
                Launch L("cov_order_kernel", stream);
                cov_order_kernel<<<(unsigned)n_ev_tiles, 256, 0, stream>>>(starts, ends, ev_off, off, n_contigs, n_events,
                                                                        ev_first.as<int32_t>(), k_max.as<OrderTile>(),
                                                                        flags.as<unsigned int>());
            }
            {
                Launch L("cov_order_scan_kernel", stream);
                cov_order_scan_kernel<<<1, 1024, 0, stream>>>(k_max.as<OrderTile>(), n_ev_tiles, flags.as<unsigned int>());
            }
        }
        MP2_CUDA(cudaGetLastError());
        const unsigned int *meas = measure ? flags.as<unsigned int>() : nullptr;
        {
            Launch L("cov_bounds_kernel", stream);
            if (fast_ver == 1)
                cov_bounds_kernel<<<(unsigned)((n_units + 127) / 128), 128, 0, stream>>>(
                    starts, ev_off, off, n_contigs, n_units, max_span, meas, e_lo.as<int64_t>(), e_hi.as<int64_t>(),
                    c_lo.as<int32_t>(), c_hi.as<int32_t>());
            else
                cov_bounds_run_kernel<<<(unsigned)((n_units + 127) / 128), 128, 0, stream>>>(
                    starts, ev_off, off, n_contigs, n_units, max_span, max_disorder, need_sorted ? 1 : 0, meas,
                    e_lo.as<int64_t>(), e_hi.as<int64_t>(), c_lo.as<int32_t>(), c_hi.as<int32_t>(), c_ev.as<int32_t>(),
                    gate, ordered ? t_wcls.as<unsigned int>() : nullptr);
        }
        MP2_CUDA(cudaGetLastError());
        if (ordered) {
            Launch L("cov_run_order_kernel", stream);
            cov_run_order_kernel<<<(unsigned)((n_units + 255) / 256), 256, 0, stream>>>(
                e_lo.as<int64_t>(), e_hi.as<int64_t>(), n_units, t_wcls.as<unsigned int>(),
                t_wcls.as<unsigned int>() + C4_WEIGHT_CLASSES, t_order.as<int32_t>(), gate);
        }
        MP2_CUDA(cudaGetLastError());
        TileArgs t;
        t.starts = starts; t.ends = ends; t.ev_off = ev_off; t.off = off;
        t.n = n_contigs; t.n_pos = n_pos; t.n_tiles = n_units;
        t.e_lo = e_lo.as<int64_t>(); t.e_hi = e_hi.as<int64_t>();
        t.c_lo = c_lo.as<int32_t>(); t.c_hi = c_hi.as<int32_t>(); t.c_ev = c_ev.as<int32_t>();
        t.n_events = n_events;
        t.rows = t_rows.as<uint32_t>(); t.meta = t_meta.as<RowMeta>();
        t.deep_list = t_deep.as<int32_t>(); t.deep_count = ctrs.as<int>() + 0;
        t.work_counter = ctrs.as<int>() + 1;
        t.order = ordered ? t_order.as<int32_t>() : nullptr;
        t.deep_row_of = t_drow.as<int32_t>(); t.deep_pool = t_dpool.as<uint32_t>(); t.deep_rows_used = ctrs.as<int>() + 6;
        t.deep_rows = deep_rows;
        t.check = measure ? flags.as<unsigned int>() : nullptr;  // tile kernel: flag[0] != 0 => not sorted
        t.span_host = max_span;
        t.span_dev = measure ? flags.as<unsigned int>() + 1 : nullptr;
        t.gate_out = gate;
        t.cov = (float *)d_cov; t.cov_stride = cov_stride; t.stats = (mp2_cov_stats *)d_stats; t.tp = tp;
        int occ = 0;
        if (fast_ver == 1) {
            MP2_TRY(cov_tile_attr_once());
            MP2_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, cov_tile_kernel, COV_THREADS, CT_SMEM));
            if (occ < 1) occ = 1;
            int64_t grid = (int64_t)num_sms() * occ;
            if (grid > n_units) grid = n_units;
            Launch L("cov_tile_kernel", stream);
            cov_tile_kernel<<<(unsigned)grid, COV_THREADS, CT_SMEM, stream>>>(t);
        } else if (fast_ver == 3) {
            MP2_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, cov_stream_kernel, 32 * C3_WPC, 0));
            if (occ < 1) occ = 1;
            int64_t grid = (int64_t)num_sms() * occ;
            if (grid * C3_WPC > n_units) grid = (n_units + C3_WPC - 1) / C3_WPC;
            Launch L("cov_stream_kernel", stream);
            cov_stream_kernel<<<(unsigned)grid, 32 * C3_WPC, 0, stream>>>(t);
        } else {
            MP2_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, cov_stream2_kernel, 32, 0));
            if (occ < 1) occ = 1;
            int64_t grid = (int64_t)num_sms() * occ;
            if (grid > n_units) grid = n_units;
            Launch L("cov_stream2_kernel", stream);
            cov_stream2_kernel<<<(unsigned)grid, 32, 0, stream>>>(t);
        }
        MP2_CUDA(cudaGetLastError());
        // The general path needs 4 bytes per position of scratch (12 GB at C3, 120 GB for a 30 Gbp sample).
        //  * scratch that is small next to the device (<= a quarter of its memory): the fallback is enqueued right
        //    behind the fast path with every kernel gated on the flag (they exit at once when it is clear) -- the
        //    call stays asynchronous and the host keeps running ahead of the device;
        //  * larger: the flag is read back first (one stream synchronisation) and the scratch is only allocated
        //    when the sample really has to be redone.
        size_t free_b = 0, total_b = 0;
        MP2_CUDA(cudaMemGetInfo(&free_b, &total_b));
        const size_t scratch = sizeof(int32_t) * (size_t)(n_pos + 4) +
                               sizeof(uint32_t) * COV_BINS * (size_t)((n_pos + COV_TILE - 1) / COV_TILE);
        if (scratch <= total_b / 4)
            return cov_general_path(starts, ends, ev_off, off, n_contigs, n_events, n_pos, ev_first.as<int32_t>(), tp,
                                    ctrs.as<int>(), gate, (float *)d_cov, cov_stride, (mp2_cov_stats *)d_stats, stream);
        int h_gate = 0;
        MP2_CUDA(cudaMemcpyAsync(&h_gate, gate, sizeof(int), cudaMemcpyDeviceToHost, stream));
        MP2_CUDA(cudaStreamSynchronize(stream));
        if (!h_gate) return MP2_OK;
    }
    return cov_general_path(starts, ends, ev_off, off, n_contigs, n_events, n_pos, ev_first.as<int32_t>(), tp,
                            ctrs.as<int>(), nullptr, (float *)d_cov, cov_stride, (mp2_cov_stats *)d_stats, stream);
}

// The two numbers mp2_cov_events_device can be vouched for with, measured exactly on the device: what a producer
// that does not know them (unlike mp2_bam_open, which measures while it decodes) calls once per sample.
extern "C" int mp2_cov_order_device(const void *d_starts, const void *d_ends, const void *d_ev_off,
                                    const void *d_contig_off, int64_t n_contigs, int64_t n_events,
                                    uint32_t *max_span, uint32_t *max_disorder, void *stream_) {
    if (n_contigs < 0 || n_events < 0) return fail(MP2_EINVAL, "negative size");
    if (!max_span || !max_disorder || !d_contig_off || !d_ev_off || (n_events > 0 && (!d_starts || !d_ends)))
        return fail(MP2_EINVAL, "NULL argument");
    MP2_TRY(require_device(-1));
    cudaStream_t stream = (cudaStream_t)stream_;
    *max_span = *max_disorder = 0;
    if (n_contigs == 0 || n_events == 0) return MP2_OK;
    const int64_t n_ev_tiles = (n_events + COV_EV_TILE - 1) / COV_EV_TILE;
    Temp flags, ev_first, tiles;
    MP2_TRY(flags.alloc(sizeof(unsigned int) * 4, stream));
    MP2_TRY(ev_first.alloc(sizeof(int32_t) * (n_ev_tiles + 1), stream));
    MP2_TRY(tiles.alloc(sizeof(OrderTile) * n_ev_tiles, stream));
    MP2_CUDA(cudaMemsetAsync(flags.p, 0, sizeof(unsigned int) * 4, stream));
    {
        Launch L("cov_map_kernel", stream);
        cov_map_kernel<<<(unsigned)((n_ev_tiles + 1 + 255) / 256), 256, 0, stream>>>(
            (const int64_t *)d_ev_off, n_contigs, COV_EV_TILE, n_ev_tiles + 1, ev_first.as<int32_t>());
    }
    {
        Launch L("cov_order_kernel", stream);
        cov_order_kernel<<<(unsigned)n_ev_tiles, 256, 0, stream>>>(
            (const uint32_t *)d_starts, (const uint32_t *)d_ends, (const int64_t *)d_ev_off, (const int64_t *)d_contig_off,
            n_contigs, n_events, ev_first.as<int32_t>(), tiles.as<OrderTile>(), flags.as<unsigned int>());
    }
    {
        Launch L("cov_order_scan_kernel", stream);
        cov_order_scan_kernel<<<1, 1024, 0, stream>>>(tiles.as<OrderTile>(), n_ev_tiles, flags.as<unsigned int>());
    }
    MP2_CUDA(cudaGetLastError());
    unsigned int h[4] = {0, 0, 0, 0};
    MP2_CUDA(cudaMemcpyAsync(h, flags.p, sizeof h, cudaMemcpyDeviceToHost, stream));
    MP2_CUDA(cudaStreamSynchronize(stream));
    *max_span = h[1];
    *max_disorder = h[2];
    return MP2_OK;
}

// ---------------- general path: difference array in HBM (any order, any span, any depth) ----------------
static int cov_general_path(const uint32_t *starts, const uint32_t *ends, const int64_t *ev_off, const int64_t *off,
                            int64_t n_contigs, int64_t n_events, int64_t n_pos, const int32_t *ev_first,
                            const TrimParams &tp, int *ctrs, const int *gate, float *d_cov, int64_t cov_stride,
                            mp2_cov_stats *d_stats, cudaStream_t stream) {
    const int64_t n_ev_tiles = (n_events + COV_EV_TILE - 1) / COV_EV_TILE;
    const int64_t n_tiles = (n_pos + COV_TILE - 1) / COV_TILE;
    const int *g = gate;  // NULL = run; else every kernel below exits at once unless *gate != 0
    Temp D, tile_sum, tile_pre, first, rows, meta, deep_list;
    MP2_TRY(D.alloc(sizeof(int32_t) * (size_t)(n_pos + 4), stream));
    MP2_TRY(tile_sum.alloc(sizeof(int32_t) * n_tiles, stream));
    MP2_TRY(tile_pre.alloc(sizeof(int32_t) * n_tiles, stream));
    MP2_TRY(first.alloc(sizeof(int32_t) * n_tiles, stream));
    MP2_TRY(rows.alloc(sizeof(uint32_t) * COV_BINS * (size_t)n_tiles, stream));
    MP2_TRY(meta.alloc(sizeof(RowMeta) * n_tiles, stream));
    MP2_TRY(deep_list.alloc(sizeof(int32_t) * n_contigs, stream));
    MP2_CUDA(cudaMemsetAsync(tile_sum.p, 0, sizeof(int32_t) * n_tiles, stream));
    MP2_CUDA(cudaMemsetAsync(meta.p, 0, sizeof(RowMeta) * n_tiles, stream));
    {
        Launch L("cov_zero_kernel", stream);
        cov_zero_kernel<<<(unsigned)(num_sms() * 8), 256, 0, stream>>>(rows.as<int4>(), (int64_t)COV_BINS / 4 * n_tiles, g);
        cov_zero_kernel<<<(unsigned)(num_sms() * 8), 256, 0, stream>>>(D.as<int4>(), (n_pos + 3) / 4, g);
    }
    MP2_CUDA(cudaGetLastError());
    {
        Launch L("cov_map_kernel", stream);
        cov_map_kernel<<<(unsigned)((n_tiles + 255) / 256), 256, 0, stream>>>(off, n_contigs, COV_TILE, n_tiles,
                                                                            first.as<int32_t>());
    }
    MP2_CUDA(cudaGetLastError());
    if (n_events > 0) {
        Launch L("cov_scatter_kernel", stream);
        cov_scatter_kernel<<<(unsigned)std::min<int64_t>(n_ev_tiles, (int64_t)num_sms() * 16), 256, 0, stream>>>(
            starts, ends, ev_off, off, n_contigs, n_events, ev_first, D.as<int32_t>(), tile_sum.as<int32_t>(), g);
    }
    MP2_CUDA(cudaGetLastError());
    {
        Launch L("cov_tile_scan_kernel", stream);
        cov_tile_scan_kernel<<<1, 1024, 0, stream>>>(tile_sum.as<int32_t>(), n_tiles, tile_pre.as<int32_t>(), g);
    }
    MP2_CUDA(cudaGetLastError());
    DepthArgs a;
    a.D = D.as<int32_t>();
    a.off = off;
    a.n = n_contigs;
    a.n_pos = n_pos;
    a.first = first.as<int32_t>();
    a.tile_pre = tile_pre.as<int32_t>();
    a.n_tiles = n_tiles;
    a.rows = rows.as<uint32_t>();
    a.meta = meta.as<RowMeta>();
    a.deep_list = deep_list.as<int32_t>();
    a.deep_count = ctrs + 3;
    a.work_counter = ctrs + 4;
    a.gate = g;
    a.cov = d_cov;
    a.cov_stride = cov_stride;
    a.stats = d_stats;
    a.tp = tp;
    int occ = 0;
    MP2_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, cov_depth_kernel, COV_THREADS, 0));
    if (occ < 1) occ = 1;
    int64_t grid = (int64_t)num_sms() * occ;
    if (grid > n_tiles) grid = n_tiles;
    {
        Launch L("cov_depth_kernel", stream);
        cov_depth_kernel<<<(unsigned)grid, COV_THREADS, 0, stream>>>(a);
    }
    MP2_CUDA(cudaGetLastError());
    {
        Launch L("cov_deep_kernel", stream);
        cov_deep_kernel<<<(unsigned)(num_sms() * 2), COV_THREADS, 0, stream>>>(a, ctrs + 5);
    }
    MP2_CUDA(cudaGetLastError());
    return MP2_OK;
}
