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
#include "mp2_common.h"

namespace mp2 {

constexpr int TNF_THREADS = 256;
constexpr int TNF_CHUNK = 65536;     // bytes of the stream owned by one work item
constexpr int TNF_SMALL_SEG = 1024;  // segments shorter than this go straight to global atomics

__constant__ uint8_t c_lut256[256];         // code -> canonical class
__constant__ uint8_t c_canon_code[MP2_NCANON];  // class -> smallest code

static int upload_tables() {
    static bool done[64] = {false};
    int dev = 0;
    MP2_CUDA(cudaGetDevice(&dev));
    if (dev < 64 && done[dev]) return MP2_OK;
    MP2_CUDA(cudaMemcpyToSymbol(c_lut256, host_lut256(), 256));
    MP2_CUDA(cudaMemcpyToSymbol(c_canon_code, host_canon_code136(), MP2_NCANON));
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

// ---- chunk -> first contig map -----------------------------------------------------------
__global__ void tnf_chunk_map_kernel(const int64_t *__restrict__ offsets, int64_t n, int mis,
                                     int64_t chunk_begin, int64_t n_chunks,
                                     int32_t *__restrict__ first_contig) {
    int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (k >= n_chunks) return;
    int64_t pos = (chunk_begin + k) * TNF_CHUNK - mis;  // stream position of the chunk start
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
__global__ void __launch_bounds__(TNF_THREADS, 4)
tnf_kernel(const uint8_t *__restrict__ bases_al, int mis, int64_t end_q,
           const int64_t *__restrict__ offsets, int64_t n,
           const int32_t *__restrict__ first_contig, int64_t chunk_begin, int64_t n_chunks,
           uint32_t *__restrict__ counts, unsigned long long *__restrict__ gc_counts,
           int *__restrict__ work_counter) {
    __shared__ uint32_t hist[256 * 32];  // [code][lane]
    __shared__ uint32_t red[256];
    __shared__ unsigned long long s_gc[2];
    __shared__ int s_chunk;

    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int warp = tid >> 5;
    uint32_t *hist_lane = hist + lane;
    char *hist_bytes = reinterpret_cast<char *>(hist);
    const uint32_t lane4 = lane * 4;

    for (int i = tid; i < 256 * 32; i += TNF_THREADS) hist[i] = 0;
    if (tid < 2) s_gc[tid] = 0;
    int par = 0;
    __syncthreads();

    for (;;) {
        if (tid == 0) s_chunk = atomicAdd(work_counter, 1);
        __syncthreads();
        const int64_t kc = s_chunk;
        __syncthreads();
        if (kc >= n_chunks) break;
        const int64_t chunk_lo_raw = (chunk_begin + kc) * TNF_CHUNK;
        const int64_t chunk_lo = chunk_lo_raw < mis ? mis : chunk_lo_raw;
        int64_t chunk_hi = chunk_lo_raw + TNF_CHUNK;
        if (chunk_hi > end_q) chunk_hi = end_q;

        for (int64_t c = first_contig[kc]; c < n; c++) {
            const int64_t o0 = offsets[c] + mis;
            if (o0 >= chunk_hi) break;
            const int64_t o1 = offsets[c + 1] + mis;
            const int64_t seg_lo = o0 > chunk_lo ? o0 : chunk_lo;
            const int64_t seg_hi = o1 < chunk_hi ? o1 : chunk_hi;
            if (seg_hi <= seg_lo) continue;

            if (seg_hi - seg_lo < TNF_SMALL_SEG) {
                // ---- tiny segment: no shared histogram, straight to the contig's row ----
                uint32_t gcl = 0;
                for (int64_t q = seg_lo + tid; q < seg_hi; q += TNF_THREADS) {
                    uint32_t c3 = code1(bases_al[q]);
                    gcl += (c3 == 1 || c3 == 2);
                    if (q - 3 >= o0 && c3 < 4) {
                        uint32_t c0 = code1(bases_al[q - 3]);
                        uint32_t c1 = code1(bases_al[q - 2]);
                        uint32_t c2 = code1(bases_al[q - 1]);
                        if ((c0 | c1 | c2) < 4) {
                            uint32_t code = (c0 << 6) | (c1 << 4) | (c2 << 2) | c3;
                            atomicAdd(&counts[c * MP2_NCANON + c_lut256[code]], 1u);
                        }
                    }
                }
                gcl = __reduce_add_sync(0xffffffffu, gcl);
                if (lane == 0 && gcl) atomicAdd(&gc_counts[c], (unsigned long long)gcl);
                continue;
            }

            // ---- regular segment: shared-memory histogram ----
            uint32_t gcl = 0;
            const int64_t g_first = seg_lo >> 4, g_last = (seg_hi - 1) >> 4;
            for (int64_t g = g_first + tid; g <= g_last; g += TNF_THREADS) {
                const int64_t P = g << 4;
                bool fast = (P - 3 >= o0) && (P >= seg_lo) && (P + 16 <= seg_hi);
                uint32_t cw = 0, prev = 0;
                if (fast) {
                    const uint4 v = __ldg(reinterpret_cast<const uint4 *>(bases_al + P));
                    const uint32_t h = __ldg(reinterpret_cast<const uint32_t *>(bases_al + P - 4));
                    uint32_t d0, d1, d2, d3, dh;
                    uint32_t p0 = decode4(v.x, d0), p1 = decode4(v.y, d1);
                    uint32_t p2 = decode4(v.z, d2), p3 = decode4(v.w, d3);
                    prev = decode4(h, dh);
                    cw = (p0 << 24) | (p1 << 16) | (p2 << 8) | p3;
                    fast = ((d0 | d1 | d2 | d3 | (dh & 0xFFFFFF00u)) == 0);
                }
                if (fast) {
                    gcl += __popc((cw ^ (cw >> 1)) & 0x55555555u);
                    // k-mer ending at base j = bits [2(15-j)+7 : 2(15-j)] of (prev:cw); its
                    // counter sits at byte offset code*128 + lane*4, so shift straight to bit 7
#pragma unroll
                    for (int j = 0; j < 16; j++) {
                        const int sh = 2 * (15 - j);
                        const uint32_t t = sh >= 7 ? __funnelshift_r(cw, prev, sh - 7)
                                                   : (cw << (7 - sh));
                        atomicAdd(reinterpret_cast<uint32_t *>(hist_bytes + ((t & 0x7F80u) | lane4)),
                                  1u);
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
                            if (run >= 4) atomicAdd(hist_lane + (kmer << 5), 1u);
                        }
                    }
                }
            }
            gcl = __reduce_add_sync(0xffffffffu, gcl);
            if (lane == 0 && gcl) atomicAdd(&s_gc[par], (unsigned long long)gcl);
            __syncthreads();
            // fold the 32 lane columns; warp w owns codes 32w..32w+31, lane l keeps code 32w+l
            uint32_t mine = 0;
#pragma unroll 4
            for (int i = 0; i < 32; i++) {
                const int code = warp * 32 + i;
                uint32_t v = hist[code * 32 + lane];
                hist[code * 32 + lane] = 0;
                v = __reduce_add_sync(0xffffffffu, v);
                if (lane == i) mine = v;
            }
            red[tid] = mine;
            __syncthreads();
            if (tid < MP2_NCANON) {
                const uint32_t a = c_canon_code[tid];
                uint32_t rc = 0, t = a;
#pragma unroll
                for (int i = 0; i < 4; i++) { rc = (rc << 2) | (3u - (t & 3u)); t >>= 2; }
                uint32_t v = red[a];
                if (rc != a) v += red[rc];
                if (v) atomicAdd(&counts[c * MP2_NCANON + tid], v);
            } else if (tid == MP2_NCANON) {
                unsigned long long v = s_gc[par];
                s_gc[par] = 0;
                if (v) atomicAdd(&gc_counts[c], v);
            }
            par ^= 1;
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
    int64_t grid = (int64_t)num_sms() * 4;
    if (grid > n_chunks) grid = n_chunks;
    {
        Launch L("tnf_kernel", stream);
        tnf_kernel<<<(unsigned)grid, TNF_THREADS, 0, stream>>>(
            d_bases - mis, mis, n_bases + mis, d_offsets, n, map.as<int32_t>(), chunk_begin,
            n_chunks, d_counts, d_gc_counts, ctr.as<int>());
    }
    MP2_CUDA(cudaGetLastError());
    return MP2_OK;
}

int64_t tnf_num_chunks(const void *d_bases, int64_t n_bases) {
    const int mis = (int)(reinterpret_cast<uintptr_t>(d_bases) & 15);
    return (n_bases + mis + TNF_CHUNK - 1) / TNF_CHUNK;
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
