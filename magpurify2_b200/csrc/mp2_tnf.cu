// Composition path: canonical 4-mer counts + GC per contig on sm_100a.
//
// Replaces count_kmers / count_canonical_fourmers / get_tnf / gc_content / get_gc
// (/root/reference/src/composition.rs:42-63, 67-74, 94-124, 128-132, 148-161).
//
// Layout in HBM: all contigs of a batch concatenated as one byte stream (1 B/base) plus
// int64 offsets[n+1].  The stream is cut into fixed 64 KiB chunks; a chunk owns the k-mers that
// END inside it, so work per CTA is independent of contig length (a 10 Mbp contig is just 160
// chunks, a chunk full of tiny contigs is a loop over segments).  Per segment (= chunk ∩ contig)
// a CTA counts into a 256-code x 32-lane shared-memory histogram: word (code, lane) lives in
// bank `lane`, so every shared atomic issued by a warp is bank-conflict free no matter which
// codes the lanes hit.  The flush folds the 32 lane columns with warp reductions, folds
// 256 codes -> 136 canonical classes and adds the non-zero classes to the contig's row with
// integer global atomics (order independent => bit exact).
#include <cstdlib>

#include "mp2_common.h"

namespace mp2 {

constexpr int TNF_WARPS = 8;
constexpr int TNF_THREADS = TNF_WARPS * 32;
constexpr int TNF_PIECE = 16384;  // bytes of the stream owned by one work item (one warp)
constexpr int TNF_UNROLL = 4;     // 16-byte groups in flight per lane

// Table words per warp: MODE 4 counts one 4-mer per base into T4[256]; MODE 5 counts one 5-mer
// per TWO bases into T5[1024] (a 5-mer b0..b4 stands for the 4-mers b0..b3 and b1..b4) and
// keeps T4[256] for the odd ones out (segment edges, groups holding non-ACGT bytes).
template <int MODE> struct TnfTab { static constexpr int WORDS = MODE == 5 ? 1280 : 256; };

__constant__ uint8_t c_lut256[256];             // code -> canonical class
__constant__ uint16_t c_canon_pair[MP2_NCANON];  // class -> smallest code | revcomp code << 8

static int upload_tables() {
    static bool done[64] = {false};
    int dev = 0;
    MP2_CUDA(cudaGetDevice(&dev));
    if (dev < 64 && done[dev]) return MP2_OK;
    MP2_CUDA(cudaMemcpyToSymbol(c_lut256, host_lut256(), 256));
    uint16_t pair[MP2_NCANON];
    const uint8_t *cc = host_canon_code136();
    for (int k = 0; k < MP2_NCANON; k++) {
        unsigned a = cc[k], t = a, rc = 0;
        for (int i = 0; i < 4; i++) { rc = (rc << 2) | (3u - (t & 3u)); t >>= 2; }
        pair[k] = (uint16_t)(a | (rc << 8));
    }
    MP2_CUDA(cudaMemcpyToSymbol(c_canon_pair, pair, sizeof pair));
    if (dev < 64) done[dev] = true;
    return MP2_OK;
}

// ---- byte decoding ---------------------------------------------------------------------
// 4 ASCII bases in one 32-bit word -> 8 bits of 2-bit codes (first base most significant).
// `diff` is zero in every byte that is one of AaCcGgTt and non-zero otherwise:
//   code  = ((b>>1) ^ (b>>2)) & 3            A=0 C=1 G=2 T=3 (bit 5, the case bit, is unused)
//   isT   = bit2 & ~bit1
//   valid <=> (b & 0xD9) == (isT ? 0x50 : 0x41)
__device__ __forceinline__ uint32_t decode4(uint32_t w, uint32_t &diff) {
    uint32_t s1 = w >> 1, s2 = w >> 2;
    uint32_t x = (s1 ^ s2) & 0x03030303u;
    uint32_t isT = s2 & ~s1 & 0x01010101u;
    diff = ((w & 0xD9D9D9D9u) ^ 0x41414141u) ^ (isT * 0x11u);
    return (x * 0x40100401u) >> 24;
}

// scalar: 0..3 for AaCcGgTt, 4 otherwise (src/composition.rs:49-55)
__device__ __forceinline__ uint32_t code1(uint32_t b) {
    uint32_t u = b & 0xDFu;
    if (u == 'A') return 0;
    if (u == 'C') return 1;
    if (u == 'G') return 2;
    if (u == 'T') return 3;
    return 4;
}

__device__ __forceinline__ void smem_inc(uint32_t saddr) {
    asm volatile("red.shared.add.u32 [%0], 1;" ::"r"(saddr) : "memory");
}

// ---- piece -> first contig map -----------------------------------------------------------
__global__ void tnf_chunk_map_kernel(const int64_t *__restrict__ offsets, int64_t n, int mis,
                                     int64_t chunk_begin, int64_t n_chunks,
                                     int32_t *__restrict__ first_contig) {
    int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (k >= n_chunks) return;
    int64_t pos = (chunk_begin + k) * TNF_PIECE - mis;  // stream position of the piece start
    // largest c in [0, n-1] with offsets[c] <= pos (0 if none)
    int64_t lo = 0, hi = n;  // search upper_bound over offsets[0..n)
    while (lo < hi) {
        int64_t mid = (lo + hi) >> 1;
        if (offsets[mid] <= pos) lo = mid + 1;
        else hi = mid;
    }
    first_contig[k] = (int32_t)(lo > 0 ? lo - 1 : 0);
}

// ---- main kernel -------------------------------------------------------------------------
// Every warp works alone: it draws a piece of the stream from a global counter, walks the
// segments (piece ∩ contig) in it, counts each segment into its own 1 KB (MODE 4) / 5 KB
// (MODE 5) shared-memory table with red.shared (same-address lanes are merged by the hardware,
// measured 13.4 increments/clk/SM on B200 for a 256-word table) and flushes the table to the
// contig's row of `counts` with integer global reductions (order independent => bit exact).
// No block-level synchronisation anywhere.
template <int MODE>
__global__ void __launch_bounds__(TNF_THREADS)
tnf_kernel(const uint8_t *__restrict__ bases_al, int mis, int64_t end_q,
           const int64_t *__restrict__ offsets, int64_t n,
           const int32_t *__restrict__ first_contig, int64_t chunk_begin, int64_t n_chunks,
           uint32_t *__restrict__ counts, unsigned long long *__restrict__ gc_counts,
           int *__restrict__ work_counter) {
    constexpr int WORDS = TnfTab<MODE>::WORDS;
    __shared__ __align__(1024) uint32_t s_tab[TNF_WARPS * WORDS];

    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    uint32_t *tab = s_tab + warp * WORDS;
    const uint32_t t5 = (uint32_t)__cvta_generic_to_shared(tab);  // T5 (MODE 5) or T4 (MODE 4)
    const uint32_t t4 = MODE == 5 ? t5 + 4096u : t5;

    for (int i = lane; i < WORDS; i += 32) tab[i] = 0;
    __syncwarp();

    for (;;) {
        int64_t kc = 0;
        if (lane == 0) kc = atomicAdd(work_counter, 1);
        kc = __shfl_sync(0xffffffffu, kc, 0);
        if (kc >= n_chunks) break;
        const int64_t piece_lo_raw = (chunk_begin + kc) * TNF_PIECE;
        const int64_t piece_lo = piece_lo_raw < mis ? mis : piece_lo_raw;
        int64_t piece_hi = piece_lo_raw + TNF_PIECE;
        if (piece_hi > end_q) piece_hi = end_q;

        for (int64_t c = first_contig[kc]; c < n; c++) {
            const int64_t o0 = offsets[c] + mis;
            if (o0 >= piece_hi) break;
            const int64_t o1 = offsets[c + 1] + mis;
            const int64_t seg_lo = o0 > piece_lo ? o0 : piece_lo;
            const int64_t seg_hi = o1 < piece_hi ? o1 : piece_hi;
            if (seg_hi <= seg_lo) continue;

            uint32_t gcl = 0;
            const int64_t g_last = (seg_hi - 1) >> 4;
            for (int64_t g0 = (seg_lo >> 4) + lane; g0 <= g_last; g0 += 32 * TNF_UNROLL) {
                uint4 v[TNF_UNROLL];
                uint32_t h[TNF_UNROLL];
                bool fast[TNF_UNROLL];
#pragma unroll
                for (int u = 0; u < TNF_UNROLL; u++) {
                    const int64_t P = (g0 + 32 * u) << 4;
                    fast[u] = (P - 3 >= o0) && (P >= seg_lo) && (P + 16 <= seg_hi);
                    if (fast[u]) {
                        v[u] = __ldg(reinterpret_cast<const uint4 *>(bases_al + P));
                        h[u] = __ldg(reinterpret_cast<const uint32_t *>(bases_al + P - 4));
                    }
                }
#pragma unroll
                for (int u = 0; u < TNF_UNROLL; u++) {
                    const int64_t P = (g0 + 32 * u) << 4;
                    if (g0 + 32 * u > g_last) break;
                    uint32_t cw = 0, prev = 0;
                    bool ok = fast[u];
                    if (ok) {
                        uint32_t d0, d1, d2, d3, dh;
                        const uint32_t p0 = decode4(v[u].x, d0), p1 = decode4(v[u].y, d1);
                        const uint32_t p2 = decode4(v[u].z, d2), p3 = decode4(v[u].w, d3);
                        prev = decode4(h[u], dh);
                        cw = (p0 << 24) | (p1 << 16) | (p2 << 8) | p3;
                        ok = ((d0 | d1 | d2 | d3 | (dh & 0xFFFFFF00u)) == 0);
                    }
                    if (ok) {
                        gcl += __popc((cw ^ (cw >> 1)) & 0x55555555u);
                        if (MODE == 5) {
                            // 5-mer ending at base j+1 (j even) = bits [2(14-j)+9 : 2(14-j)] of
                            // prev:cw; its counter is word index*4 bytes into T5
#pragma unroll
                            for (int j = 0; j < 16; j += 2) {
                                const int sh = 2 * (14 - j);
                                const uint32_t t = sh >= 2 ? __funnelshift_r(cw, prev, sh - 2) : (cw << 2);
                                smem_inc(t5 + (t & 0xFFCu));
                            }
                        } else {
#pragma unroll
                            for (int j = 0; j < 16; j++) {
                                const int sh = 2 * (15 - j);
                                const uint32_t t = sh >= 2 ? __funnelshift_r(cw, prev, sh - 2) : (cw << 2);
                                smem_inc(t4 + (t & 0x3FCu));
                            }
                        }
                    } else {
                        // boundary group or a group holding non-ACGT bytes: rolling scalar walk
                        uint32_t kmer = 0, run = 0;
                        for (int i = -3; i < 16; i++) {
                            const int64_t q = P + i;
                            if (q >= seg_hi) break;
                            if (q < o0) { run = 0; continue; }
                            const uint32_t cd = code1(bases_al[q]);
                            if (cd > 3) run = 0;
                            else { kmer = ((kmer << 2) | cd) & 0xFFu; run++; }
                            if (i >= 0 && q >= seg_lo) {
                                gcl += (cd == 1 || cd == 2);
                                if (run >= 4) smem_inc(t4 + (kmer << 2));
                            }
                        }
                    }
                }
            }
            gcl = __reduce_add_sync(0xffffffffu, gcl);
            if (lane == 0 && gcl) atomicAdd(&gc_counts[c], (unsigned long long)gcl);
            __syncwarp();
            // ---- flush: fold 256 codes -> 136 canonical classes, add to the contig's row ----
#pragma unroll
            for (int k = lane; k < MP2_NCANON + 24; k += 32) {
                if (k < MP2_NCANON) {
                    const uint32_t pr = c_canon_pair[k];
                    const uint32_t a = pr & 0xFFu, rc = pr >> 8;
                    uint32_t val = tab[(MODE == 5 ? 1024 : 0) + a];
                    if (MODE == 5) {
                        const uint4 q = *reinterpret_cast<const uint4 *>(tab + (a << 2));
                        val += q.x + q.y + q.z + q.w + tab[a] + tab[256 + a] + tab[512 + a] + tab[768 + a];
                    }
                    if (rc != a) {
                        val += tab[(MODE == 5 ? 1024 : 0) + rc];
                        if (MODE == 5) {
                            const uint4 q = *reinterpret_cast<const uint4 *>(tab + (rc << 2));
                            val += q.x + q.y + q.z + q.w + tab[rc] + tab[256 + rc] + tab[512 + rc] + tab[768 + rc];
                        }
                    }
                    if (val) atomicAdd(&counts[c * MP2_NCANON + k], val);
                }
            }
            __syncwarp();
#pragma unroll
            for (int i = 0; i < WORDS / 128; i++)
                *reinterpret_cast<uint4 *>(tab + (i * 32 + lane) * 4) = make_uint4(0, 0, 0, 0);
            __syncwarp();
        }
    }
}

// ---- finalisation --------------------------------------------------------------------------
// relative frequencies (src/composition.rs:111-118): u32 -> f32, row sum in f32, divide,
// NaN -> 0.  The f32 row sum equals the exact integer sum while it is < 2^24; above that the
// lane-0 fallback reproduces ndarray's 8-way unrolled fold order (see oracle/mp2_oracle.c).
__global__ void tnf_rel_kernel(const uint32_t *__restrict__ counts, int64_t n,
                               float *__restrict__ rel) {
    const int lane = threadIdx.x & 31;
    const int64_t row = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    if (row >= n) return;
    const uint32_t *r = counts + row * MP2_NCANON;
    uint32_t v[5];
    unsigned long long s = 0;
#pragma unroll
    for (int i = 0; i < 5; i++) {
        int j = lane + 32 * i;
        v[i] = j < MP2_NCANON ? r[j] : 0u;
        s += v[i];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    float sum;
    if (s < (1ull << 24)) {
        sum = (float)s;
    } else {
        float acc = 0.0f;
        if (lane == 0) {
            float p[8] = {0, 0, 0, 0, 0, 0, 0, 0};
            for (int i = 0; i < MP2_NCANON; i += 8)
                for (int j = 0; j < 8; j++) p[j] = __fadd_rn(p[j], (float)r[i + j]);
            acc = __fadd_rn(acc, __fadd_rn(p[0], p[4]));
            acc = __fadd_rn(acc, __fadd_rn(p[1], p[5]));
            acc = __fadd_rn(acc, __fadd_rn(p[2], p[6]));
            acc = __fadd_rn(acc, __fadd_rn(p[3], p[7]));
        }
        sum = __shfl_sync(0xffffffffu, acc, 0);
    }
#pragma unroll
    for (int i = 0; i < 5; i++) {
        int j = lane + 32 * i;
        if (j < MP2_NCANON) {
            float q = __fdiv_rn((float)v[i], sum);
            rel[row * MP2_NCANON + j] = (q != q) ? 0.0f : q;
        }
    }
}

// GC fraction (src/composition.rs:128-132): (f32)gc / (f32)len; 0/0 stays NaN.
__global__ void tnf_gc_kernel(const unsigned long long *__restrict__ gc_counts,
                              const int64_t *__restrict__ offsets, int64_t n,
                              float *__restrict__ gc) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    unsigned long long len = (unsigned long long)(offsets[i + 1] - offsets[i]);
    gc[i] = __fdiv_rn(__ull2float_rn(gc_counts[i]), __ull2float_rn(len));
}

// MP2_TNF_MODE=4|5 selects the counting scheme (development switch; default below).
static int tnf_mode() {
    static int mode = 0;
    if (!mode) {
        const char *e = getenv("MP2_TNF_MODE");
        mode = (e && e[0] == '4') ? 4 : 5;
    }
    return mode;
}

// Enqueue the counting kernels for chunks [chunk_begin, chunk_end) of the stream.
// d_counts / d_gc_counts must have been zeroed (once) before the first range.
int tnf_launch_range(const uint8_t *d_bases, const int64_t *d_offsets, int64_t n, int64_t n_bases,
                     int64_t chunk_begin, int64_t chunk_end, uint32_t *d_counts,
                     unsigned long long *d_gc_counts, cudaStream_t stream) {
    MP2_TRY(upload_tables());
    const int mis = (int)(reinterpret_cast<uintptr_t>(d_bases) & 15);
    const int64_t n_chunks = chunk_end - chunk_begin;
    if (n_chunks <= 0) return MP2_OK;
    Temp map, ctr;
    MP2_TRY(map.alloc(sizeof(int32_t) * n_chunks, stream));
    MP2_TRY(ctr.alloc(sizeof(int), stream));
    MP2_CUDA(cudaMemsetAsync(ctr.p, 0, sizeof(int), stream));
    {
        Launch L("tnf_chunk_map_kernel", stream);
        tnf_chunk_map_kernel<<<(unsigned)((n_chunks + 255) / 256), 256, 0, stream>>>(
            d_offsets, n, mis, chunk_begin, n_chunks, map.as<int32_t>());
    }
    MP2_CUDA(cudaGetLastError());
    const int mode = tnf_mode();
    int occ = 0;
    if (mode == 5)
        MP2_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, tnf_kernel<5>, TNF_THREADS, 0));
    else
        MP2_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, tnf_kernel<4>, TNF_THREADS, 0));
    if (occ < 1) occ = 1;
    int64_t grid = (int64_t)num_sms() * occ;
    const int64_t need = (n_chunks + TNF_WARPS - 1) / TNF_WARPS;
    if (grid > need) grid = need;
    {
        Launch L("tnf_kernel", stream);
        if (mode == 5)
            tnf_kernel<5><<<(unsigned)grid, TNF_THREADS, 0, stream>>>(
                d_bases - mis, mis, n_bases + mis, d_offsets, n, map.as<int32_t>(), chunk_begin,
                n_chunks, d_counts, d_gc_counts, ctr.as<int>());
        else
            tnf_kernel<4><<<(unsigned)grid, TNF_THREADS, 0, stream>>>(
                d_bases - mis, mis, n_bases + mis, d_offsets, n, map.as<int32_t>(), chunk_begin,
                n_chunks, d_counts, d_gc_counts, ctr.as<int>());
    }
    MP2_CUDA(cudaGetLastError());
    return MP2_OK;
}

// Pieces wholly inside the first `done` bytes of the stream (host staging pipelines on this).
int64_t tnf_chunks_ready(const void *d_bases, int64_t done) {
    const int mis = (int)(reinterpret_cast<uintptr_t>(d_bases) & 15);
    return (done + mis) / TNF_PIECE;
}

int64_t tnf_num_chunks(const void *d_bases, int64_t n_bases) {
    const int mis = (int)(reinterpret_cast<uintptr_t>(d_bases) & 15);
    return (n_bases + mis + TNF_PIECE - 1) / TNF_PIECE;
}

int tnf_finalize(const uint32_t *d_counts, const unsigned long long *d_gc_counts,
                 const int64_t *d_offsets, int64_t n, float *d_rel, float *d_gc,
                 cudaStream_t stream) {
    if (n <= 0) return MP2_OK;
    if (d_rel) {
        int64_t threads = n * 32;
        {
            Launch L("tnf_rel_kernel", stream);
            tnf_rel_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, stream>>>(d_counts, n, d_rel);
        }
        MP2_CUDA(cudaGetLastError());
    }
    if (d_gc) {
        {
            Launch L("tnf_gc_kernel", stream);
            tnf_gc_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(d_gc_counts, d_offsets,
                                                                          n, d_gc);
        }
        MP2_CUDA(cudaGetLastError());
    }
    return MP2_OK;
}

}  // namespace mp2

extern "C" int mp2_tnf_device(const void *d_bases, const void *d_offsets, int64_t n,
                              int64_t n_bases, void *d_counts, void *d_rel, void *d_gc,
                              void *stream_) {
    using namespace mp2;
    if (n < 0 || n_bases < 0) return fail(MP2_EINVAL, "negative size");
    if (n > INT32_MAX) return fail(MP2_EINVAL, "too many sequences in one call (%lld)", (long long)n);
    if (!d_offsets || !d_counts || (n_bases > 0 && !d_bases))
        return fail(MP2_EINVAL, "d_bases, d_offsets and d_counts are required");
    MP2_TRY(require_device(-1));
    cudaStream_t stream = (cudaStream_t)stream_;
    if (n == 0) return MP2_OK;
    Temp gcc;
    MP2_TRY(gcc.alloc(sizeof(unsigned long long) * n, stream));
    MP2_CUDA(cudaMemsetAsync(d_counts, 0, sizeof(uint32_t) * MP2_NCANON * n, stream));
    MP2_CUDA(cudaMemsetAsync(gcc.p, 0, sizeof(unsigned long long) * n, stream));
    MP2_TRY(tnf_launch_range((const uint8_t *)d_bases, (const int64_t *)d_offsets, n, n_bases, 0,
                             tnf_num_chunks(d_bases, n_bases), (uint32_t *)d_counts,
                             gcc.as<unsigned long long>(), stream));
    MP2_TRY(tnf_finalize((const uint32_t *)d_counts, gcc.as<unsigned long long>(),
                         (const int64_t *)d_offsets, n, (float *)d_rel, (float *)d_gc, stream));
    return MP2_OK;
}
