// Composition path: canonical 4-mer counts + GC per contig on sm_100a.
//
// Replaces count_kmers / count_canonical_fourmers / get_tnf / gc_content / get_gc
// (/root/reference/src/composition.rs:42-63, 67-74, 94-124, 128-132, 148-161).
//
// Layout in HBM: all contigs of a batch concatenated as one byte stream (1 B/base) plus
// int64 offsets[n+1].  The stream is cut into pieces of nominally 16 / 32 KiB whose boundaries snap to contig starts;
// a piece owns the k-mers that END inside it, so work per warp is independent of contig length (a 10 Mbp contig is
// just 320 pieces, a piece full of tiny contigs is a loop over segments).  A warp is a CTA: per segment
// (= piece ∩ contig) it counts into its own shared-memory table and flushes the table into the contig's row -- stored
// when the segment is the whole contig, integer global reductions otherwise (order independent => bit exact).
//
// What bounds the kernel is the shared-memory data pipe (1 wavefront/clk/SM): a warp-wide red.shared on 32 random
// words costs ~3.8 wavefronts (worst bank load), so the design minimises wavefronts: one increment per TWO bases
// (5-mer table), a flush that reads the table once with conflict-free LDS.128 and gathers the 136 classes from a
// swizzled home, one shuffle and one vote per tile of four rows, rows that hold interior groups only (the few partial
// groups of a segment are counted together, out of the streaming loop).
#include <climits>
#include <cstdlib>
#include <mutex>

#include "mp2_common.h"

namespace mp2 {

constexpr int TNF_THREADS = 32;   // one warp per CTA: the warp's table sits at a fixed shared address
// Bytes of the stream owned by one work item (one warp) = 1 << shift.  Every segment costs a table flush
// and two edge rows, so large inputs use 32 KiB pieces (fewer piece-boundary segments); small inputs
// keep 16 KiB so that there are enough pieces for all warps.  MP2_TNF_PIECE_SHIFT=14|15|16 overrides.
static int tnf_piece_shift(int64_t n_bases) {
    static int forced = -1;
    if (forced < 0) {
        const char *e = getenv("MP2_TNF_PIECE_SHIFT");
        forced = (e && atoi(e) >= 14 && atoi(e) <= 16) ? atoi(e) : 0;
    }
    if (forced) return forced;
    return n_bases >= ((int64_t)1 << 28) ? 15 : 14;
}
#ifndef MP2_TNF_UNROLL
#define MP2_TNF_UNROLL 4
#endif
constexpr int TNF_UNROLL = MP2_TNF_UNROLL;  // 16-byte groups in flight per lane
static_assert(TNF_UNROLL == 4, "the halos of a tile travel in one 32-bit shuffle, 7 bits each (mask 0x08102040 = their bits 6)");

// Table words per warp: MODE 4 counts one 4-mer per base into T4[256]; MODE 5 counts one 5-mer
// per TWO bases into T5[1024] (a 5-mer b0..b4 stands for the 4-mers b0..b3 and b1..b4) and
// keeps T4[256] for the odd ones out (segment edges, groups holding non-ACGT bytes).
template <int MODE> struct TnfTab {
    static constexpr int T5W = MODE == 5 ? 1024 : 0;
    static constexpr int WORDS = T5W + 256;
};

#ifndef MP2_TNF_TMAZERO
#define MP2_TNF_TMAZERO 1
#endif
// -DMP2_TNF_BOUNDS=1: every global load of tnf_kernel and every computed shared-memory / row index is checked and a
// violation traps (compute-sanitizer is not available on the pool the kernels are developed on; the parity tests are
// run once per change with this build: tools/gpu_r2_bounds.sh).
#ifndef MP2_TNF_BOUNDS
#define MP2_TNF_BOUNDS 0
#endif
#if MP2_TNF_BOUNDS
#define TNF_CHECK(cond) do { if (!(cond)) __trap(); } while (0)
#else
#define TNF_CHECK(cond) do { } while (0)
#endif
// MODE 5 flush: the 4 KB 5-mer table is cleared by the bulk-copy engine (cp.async.bulk from a page of zeros in
// L2) instead of 8 STS.128 per lane: the kernel is bound by the LSU data pipe (1 wavefront per clock), the copy
// engine writes shared memory without going through it, and the copy runs while the warp folds the 4-mer sums
// into the 136 classes and sends them to the contig's row.
__device__ __align__(128) uint32_t g_zero_page[1024];

__constant__ uint8_t c_lut256[256];             // code -> canonical class
__constant__ uint16_t c_canon_pair[MP2_NCANON];  // class -> smallest code | revcomp code << 8
// MODE 5 flush: the 4-mer count of code c waits for the class gather at word tnf_swz(c) of the T4 region, not at word c.
// Read at word c, the 32 classes of one gather instruction cost 11 (smallest codes) + 24 (their reverse complements,
// which differ in the TOP bits and so share banks) wavefronts per flush; with this XOR swizzle -- linear over GF(2) in
// c >> 2, so that the four codes of a uint4 stay one uint4, found by search -- 17, and the two STS.128 of the sums
// stay conflict-free.  c_canon_pair5 holds the swizzled words.
__host__ __device__ constexpr uint32_t tnf_swz_g(uint32_t v) {  // v = c >> 2 (6 bits) -> 8-bit XOR mask
    return ((v & 1u) ? 111u : 0u) ^ ((v & 2u) ? 79u : 0u) ^ ((v & 4u) ? 199u : 0u) ^ ((v & 8u) ? 226u : 0u) ^
           ((v & 16u) ? 31u : 0u) ^ ((v & 32u) ? 177u : 0u);
}
__host__ __device__ constexpr uint32_t tnf_swz(uint32_t c) { return c ^ tnf_swz_g(c >> 2); }
__constant__ uint16_t c_canon_pair5[MP2_NCANON];

static int upload_tables() {
    static std::mutex mu;  // callers may be threads driving different GPUs
    static bool done[64] = {false};
    int dev = 0;
    MP2_CUDA(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lk(mu);
    if (dev >= 0 && dev < 64 && done[dev]) return MP2_OK;
    MP2_CUDA(cudaMemcpyToSymbol(c_lut256, host_lut256(), 256));
    uint16_t pair[MP2_NCANON];
    const uint8_t *cc = host_canon_code136();
    for (int k = 0; k < MP2_NCANON; k++) {
        unsigned a = cc[k], t = a, rc = 0;
        for (int i = 0; i < 4; i++) { rc = (rc << 2) | (3u - (t & 3u)); t >>= 2; }
        pair[k] = (uint16_t)(a | (rc << 8));
    }
    MP2_CUDA(cudaMemcpyToSymbol(c_canon_pair, pair, sizeof pair));
    uint16_t pair5[MP2_NCANON];
    for (int k = 0; k < MP2_NCANON; k++) pair5[k] = (uint16_t)(tnf_swz(pair[k] & 0xFFu) | (tnf_swz(pair[k] >> 8) << 8));
    MP2_CUDA(cudaMemcpyToSymbol(c_canon_pair5, pair5, sizeof pair5));
    if (dev < 64) done[dev] = true;
    return MP2_OK;
}

// ---- byte decoding ---------------------------------------------------------------------
// 4 ASCII bases in one 32-bit word -> 8 bits of 2-bit codes (first base most significant).
//   code  = ((b>>1) ^ (b>>2)) & 3            A=0 C=1 G=2 T=3 (bit 5, the case bit, is unused)
//   isT   = bit2 & ~bit1
//   valid <=> (b & 0xD9) == (isT ? 0x50 : 0x41)
// `sig` is 0x41414141 exactly when all four bytes are one of AaCcGgTt.
__device__ __forceinline__ uint32_t decode4_hi(uint32_t w, uint32_t &sig) {
    const uint32_t s1 = w >> 1, s2 = w >> 2;
    const uint32_t x = (s1 ^ s2) & 0x03030303u;
    const uint32_t isT = s2 & ~s1 & 0x01010101u;
    sig = (w & 0xD9D9D9D9u) ^ (isT * 0x11u);
    return x * 0x40100401u;  // the four codes are in bits 31:24
}
__device__ __forceinline__ uint32_t decode4(uint32_t w, uint32_t &sig) { return decode4_hi(w, sig) >> 24; }

// 4-bit mask (first byte most significant) of the bytes of `sig` that differ from 0x41
__device__ __forceinline__ uint32_t bad4(uint32_t sig) {
    const uint32_t d = sig ^ 0x41414141u;
    const uint32_t nz = (((d & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | d) & 0x80808080u;
    return ((nz >> 7) * 0x08040201u >> 24) & 0xFu;
}

__device__ __forceinline__ void smem_inc(uint32_t saddr) {
    asm volatile("red.shared.add.u32 [%0], 1;" ::"r"(saddr) : "memory");
}

// ---- piece map ---------------------------------------------------------------------------------
// Piece k nominally starts at aligned-stream position k << piece_shift.  With `snap`, a boundary moves
// forward to the first contig start within one piece of it: the piece before it then ends with a whole
// contig and the segment that would have straddled the boundary (one more table flush) does not exist.
// adj[k] = boundary - (k << piece_shift) in [0, piece); first_contig[k] = contig holding the boundary.
// Entries 0..n_chunks are written (the kernel needs the boundary after its last piece).
__global__ void tnf_chunk_map_kernel(const int64_t *__restrict__ offsets, int64_t n, int mis, int piece_shift,
                                     int snap, int64_t chunk_begin, int64_t n_chunks,
                                     int32_t *__restrict__ first_contig, int32_t *__restrict__ adj) {
    int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (k > n_chunks) return;
    const int64_t kabs = chunk_begin + k;
    int64_t pos = (kabs << piece_shift) - mis;  // stream position of the nominal boundary
    int32_t d = 0;
    if (snap && kabs > 0) {
        int64_t lo = 0, hi = n;  // lower_bound: first c with offsets[c] >= pos
        while (lo < hi) {
            const int64_t mid = (lo + hi) >> 1;
            if (offsets[mid] < pos) lo = mid + 1;
            else hi = mid;
        }
        if (lo < n) {
            const int64_t gap = offsets[lo] - pos;
            if (gap > 0 && gap < ((int64_t)1 << piece_shift)) d = (int32_t)gap;
        }
    }
    pos += d;
    // largest c in [0, n-1] with offsets[c] <= pos (0 if none)
    int64_t lo = 0, hi = n;  // search upper_bound over offsets[0..n)
    while (lo < hi) {
        int64_t mid = (lo + hi) >> 1;
        if (offsets[mid] <= pos) lo = mid + 1;
        else hi = mid;
    }
    first_contig[k] = (int32_t)(lo > 0 ? lo - 1 : 0);
    adj[k] = d;
}

// Fold the warp's table into the contig's row: 256 codes -> 136 canonical classes, integer global
// reductions, then zero the table.  MODE 4: T4 is the 4-mer table already.
__device__ __forceinline__ void tnf_flush5(uint32_t *tab, int lane, uint32_t *__restrict__ row, uint32_t &phase, bool whole);

template <int MODE>
// `whole` (warp-uniform): the segment is the WHOLE contig, nobody else adds to its row: the 136 counts are stored
// (zeros too) and the row need not have been cleared.
__device__ __forceinline__ void tnf_flush(uint32_t *tab, int lane, uint32_t *__restrict__ row, uint32_t &phase, bool whole) {
    if (MODE == 5) {
        tnf_flush5(tab, lane, row, phase, whole);
        return;
    }
    __syncwarp();
#pragma unroll
    for (int k = lane; k < MP2_NCANON + 24; k += 32) {
        if (k < MP2_NCANON) {
            const uint32_t pr = c_canon_pair[k];
            const uint32_t a = pr & 0xFFu, rc = pr >> 8;
            uint32_t val = tab[a];
            if (rc != a) val += tab[rc];
            if (whole) row[k] = val;
            else if (val) atomicAdd(&row[k], val);
        }
    }
    __syncwarp();
    *reinterpret_cast<uint4 *>(tab + lane * 4) = make_uint4(0, 0, 0, 0);
    *reinterpret_cast<uint4 *>(tab + 128 + lane * 4) = make_uint4(0, 0, 0, 0);
    __syncwarp();
}


// MODE 5 flush.  count4[c] = T4[c] + sum_b T5[c.b] + sum_b T5[b.c]; every table read below is a
// bank-conflict-free LDS.128 (a quarter warp reads 8 distinct 16-byte bank groups) and T5 is read once:
//   * lane L owns the codes 8L..8L+7.  Their suffix marginals are the words b*256 + 8L..8L+7 for b = 0..3 and
//     T4[8L..8L+7]: two uint4s each, read in the order h ^ ((L>>2)&1) and added element-wise;
//   * the same eight uint4s, summed horizontally, are the prefix marginals of the codes 64b + 2L + h, which belong
//     to other lanes: they are handed over through the T4 region (four conflict-free 64-bit stores, two LDS.128).
// The 4-mer counts end up in the T4 region at the address of their code; the 136 classes are gathered from there.
// (Until round 2 the prefix marginals were a second pass over T5: 32 more wavefronts per flush on the pipe that
// bounds the kernel.)
__shared__ __align__(8) unsigned long long tnf_bar;   // completion of the table's bulk clear (one warp per CTA)

__device__ __forceinline__ void tnf_bar_init(int lane) {
#if MP2_TNF_TMAZERO
    if (lane == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"((uint32_t)__cvta_generic_to_shared(&tnf_bar)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
#endif
}

__device__ __forceinline__ void tnf_flush5(uint32_t *tab, int lane, uint32_t *__restrict__ row, uint32_t &phase, bool whole) {
    const uint4 *t5v = reinterpret_cast<const uint4 *>(tab);
    uint4 *t4v = reinterpret_cast<uint4 *>(tab + 1024);
    // T5 is read ONCE: the uint4 at word b*256 + 8L + 4h holds T5[b . c] for the four codes c = 8L + 4h + 0..3 (added
    // element-wise over b: their suffix marginals) and, summed horizontally, is the prefix marginal of code
    // 64b + 2L + h.  The eight prefix sums go through the T4 region (its own contents are in registers by then) to
    // the lanes that own those codes.
    const int hx = (lane >> 2) & 1;
    __syncwarp();
    uint4 sum[2];
    uint32_t ps[2][4];
#pragma unroll
    for (int hp = 0; hp < 2; hp++) {
        const int h = hp ^ hx;
        sum[hp] = t4v[2 * lane + h];
#pragma unroll
        for (int b = 0; b < 4; b++) {
            const uint4 q = t5v[b * 64 + 2 * lane + h];
            sum[hp].x += q.x; sum[hp].y += q.y; sum[hp].z += q.z; sum[hp].w += q.w;
            ps[hp][b] = (q.x + q.y) + (q.z + q.w);
        }
    }
    __syncwarp();  // every lane has read its T4 words: the region is free
#pragma unroll
    for (int b = 0; b < 4; b++)  // codes 64b + 2L (h = 0) and 64b + 2L + 1 (h = 1): one conflict-free 64-bit store
        *reinterpret_cast<uint2 *>(tab + 1024 + 64 * b + 2 * lane) =
            make_uint2(hx ? ps[1][b] : ps[0][b], hx ? ps[0][b] : ps[1][b]);
    __syncwarp();
#pragma unroll
    for (int hp = 0; hp < 2; hp++) {
        const int h = hp ^ hx;
        const uint4 p = t4v[2 * lane + h];
        sum[hp].x += p.x; sum[hp].y += p.y; sum[hp].z += p.z; sum[hp].w += p.w;
    }
    __syncwarp();  // every lane has read its prefix sums: the region takes the 4-mer counts, swizzled (tnf_swz)
    const uint32_t g_lane = tnf_swz_g(2u * (uint32_t)lane);
#pragma unroll
    for (int hp = 0; hp < 2; hp++) {
        const int h = hp ^ hx;
        const uint32_t g = g_lane ^ (h ? tnf_swz_g(1u) : 0u);  // linear: g(2L + h) = g(2L) ^ g(h)
        uint4 o = sum[hp];                                      // element e goes to position e ^ (g & 3)
        if (g & 1u) { uint32_t t = o.x; o.x = o.y; o.y = t; t = o.z; o.z = o.w; o.w = t; }
        if (g & 2u) { uint32_t t = o.x; o.x = o.z; o.z = t; t = o.y; o.y = o.w; o.w = t; }
        TNF_CHECK(((2u * (uint32_t)lane + (uint32_t)h) ^ (g >> 2)) < 64u);
        t4v[(2u * (uint32_t)lane + (uint32_t)h) ^ (g >> 2)] = o;
    }
    __syncwarp();  // every lane is done reading T5; the 4-mer counts are visible in T4
#if MP2_TNF_TMAZERO
    const uint32_t bar = (uint32_t)__cvta_generic_to_shared(&tnf_bar);
    if (lane == 0) {
        // generic-proxy reads of T5 (above) before the async-proxy writes of the copy engine
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(4096u) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                         (uint32_t)__cvta_generic_to_shared(tab)),
                     "l"(g_zero_page), "r"(4096u), "r"(bar)
                     : "memory");
    }
#else
#pragma unroll
    for (int i = 0; i < 8; i++) *reinterpret_cast<uint4 *>(tab + (i * 32 + lane) * 4) = make_uint4(0, 0, 0, 0);
#endif
#pragma unroll
    for (int k = lane; k < MP2_NCANON + 24; k += 32) {
        if (k < MP2_NCANON) {
            const uint32_t pr = c_canon_pair5[k];  // words of the two codes (equal for a palindrome)
            const uint32_t a = pr & 0xFFu, rc = pr >> 8;
            uint32_t val = tab[1024 + a];
            if (rc != a) val += tab[1024 + rc];
            if (whole) row[k] = val;
            else if (val) atomicAdd(&row[k], val);
        }
    }
    __syncwarp();
    t4v[lane] = make_uint4(0, 0, 0, 0);
    t4v[32 + lane] = make_uint4(0, 0, 0, 0);
#if MP2_TNF_TMAZERO
    {   // the table is clear when the copy has landed
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "TNF_WAIT:\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
            "@p bra TNF_DONE;\n\t"
            "bra TNF_WAIT;\n\t"
            "TNF_DONE:\n\t}" ::"r"(bar), "r"(phase)
            : "memory");
        phase ^= 1u;
    }
#else
    (void)phase;
#endif
    __syncwarp();
}

// Vector path of one group of 16 valid bases whose 3 halo bases are valid too: 8 (MODE 5) or 16
// (MODE 4) increments.  halo = codes of the 3 bases before the group in bits 5:0.
template <int MODE>
__device__ __forceinline__ void tnf_count16(uint32_t cw, uint32_t halo, uint32_t t5, uint32_t t4) {
    if (MODE == 5) {
        // 5-mer ending at base j+1 (j even) = bits [2(14-j)+9 : 2(14-j)] of halo:cw; its counter is
        // word index*4 bytes into T5.  The two windows that reach into the halo take a funnel shift, the five
        // that lie inside cw a plain shift, the last one a left shift.
        smem_inc(t5 + (__funnelshift_r(cw, halo, 26) & 0xFFCu));
        smem_inc(t5 + (__funnelshift_r(cw, halo, 22) & 0xFFCu));
        smem_inc(t5 + ((cw >> 18) & 0xFFCu));
        smem_inc(t5 + ((cw >> 14) & 0xFFCu));
        smem_inc(t5 + ((cw >> 10) & 0xFFCu));
        smem_inc(t5 + ((cw >> 6) & 0xFFCu));
        smem_inc(t5 + ((cw >> 2) & 0xFFCu));
        smem_inc(t5 + ((cw << 2) & 0xFFCu));
    } else {
#pragma unroll
        for (int j = 0; j < 16; j++) {
            const int sh = 2 * (15 - j);
            const uint32_t t = sh >= 2 ? __funnelshift_r(cw, halo, sh - 2) : (cw << 2);
            smem_inc(t4 + (t & 0x3FCu));
        }
    }
}

// Masked path of one group: W has bit (18-i) set when position P-3+i is a valid base of this contig
// below seg_hi (i = 0..2 halo, 3..18 the group); own has bit (15-j) set when base j is >= seg_lo.
// K gets bit (15-j) set when the 4-mer ENDING at base j counts; returns the number of G/C bases owned.
__device__ __forceinline__ uint32_t tnf_masked_prep(uint32_t cw, uint32_t W, uint32_t own, uint32_t &K) {
    K = W & (W >> 1) & (W >> 2) & (W >> 3) & own;
    uint32_t sp = W & own;  // spread bit b -> bit 2b
    sp = (sp | (sp << 8)) & 0x00FF00FFu;
    sp = (sp | (sp << 4)) & 0x0F0F0F0Fu;
    sp = (sp | (sp << 2)) & 0x33333333u;
    sp = (sp | (sp << 1)) & 0x55555555u;
    return __popc((cw ^ (cw >> 1)) & sp);
}

// Counts the single 4-mers of K into T4, one lane alone (packed kernel).
__device__ __forceinline__ uint32_t tnf_masked16(uint32_t cw, uint32_t ph, uint32_t W, uint32_t own, uint32_t t4) {
    uint32_t K;
    const uint32_t gcn = tnf_masked_prep(cw, W, own, K);
#pragma unroll
    for (int b = 0; b < 16; b++) {
        if (K & (1u << b)) {
            const uint32_t t = b ? __funnelshift_r(cw, ph, 2 * b - 2) : (cw << 2);
            smem_inc(t4 + (t & 0x3FCu));
        }
    }
    return gcn;
}

// range masks shared by the ASCII and the packed kernels (positions relative to the piece)
__device__ __forceinline__ void tnf_range_masks(int P, int o0, int seg_lo, int seg_hi, uint32_t &W, uint32_t &own) {
    const int lo_i = o0 - (P - 3);       // positions i < lo_i precede the contig
    if (lo_i > 0) W &= lo_i >= 19 ? 0u : (0x7FFFFu >> lo_i);
    const int hi_i = seg_hi - (P - 3);   // positions i >= hi_i are beyond the segment
    if (hi_i < 19) W &= ~((1u << (19 - hi_i)) - 1u);
    const int slo_j = seg_lo - P;
    own = slo_j > 0 ? (slo_j >= 16 ? 0u : (0xFFFFu >> slo_j)) : 0xFFFFu;
}

// ---- main kernel -------------------------------------------------------------------------
// Every warp (= CTA) works alone: it draws a 16 KiB piece of the stream from a global counter,
// walks the segments (piece ∩ contig) in it, counts each segment into its own shared-memory table
// with red.shared (ATOMS.POPC.INC: same-address lanes are merged by the hardware) and flushes
// the table to the contig's row of `counts`.  No block-level synchronisation anywhere.
//
// A lane owns 16-byte groups, interleaved across the warp so every LDG.128 is a fully coalesced
// 512-byte request; a ROW is the 32 groups of one such request.  The 3-base halo of a group comes
// from the neighbouring lane by shuffle.  A row whose 32 groups are all clean (16 valid bases, valid
// halo, inside the segment) takes the vector path with no divergence at all (MODE 5: 8 increments
// for 16 bases); any other row goes through tnf_edge_row, which is out of line so that the
// streaming loop stays small.

// The shared-memory table of the warp.  One warp per CTA => a fixed shared address.
__shared__ __align__(16) uint32_t tnf_tab[TnfTab<5>::WORDS];

// The last, partial group of the whole stream, byte by byte (bytes beyond the buffer read as 0).
__device__ __noinline__ uint4 tnf_load_tail(const uint8_t *__restrict__ pb, int P, int safe_hi) {
    uint32_t w[4] = {0, 0, 0, 0};
    for (int i = 0; i < 16 && P + i < safe_hi; i++) w[i >> 2] |= (uint32_t)pb[P + i] << (8 * (i & 3));
    return make_uint4(w[0], w[1], w[2], w[3]);
}

// Exact per-base masks of ONE group (16 bases at piece-relative position 16 g, signatures s0..s3 of its four words, codes
// cw): K gets bit (15-j) set when the 4-mer ENDING at base j counts for this segment, ph the codes of the 4 bytes before
// the group; returns the number of G/C bases of the group that belong to the segment.
__device__ __forceinline__ uint32_t tnf_group_masks(uint32_t s0, uint32_t s1, uint32_t s2, uint32_t s3, uint32_t cw, int g,
                                                    int o0, int seg_lo, int seg_hi, const uint8_t *__restrict__ pb,
                                                    bool head_ok, uint32_t &K, uint32_t &ph) {
    const int P = g << 4;
    uint32_t hw = 0;  // halo bytes straight from memory (positions P-4..P-1)
    if (P >= 4 || head_ok) hw = __ldg(reinterpret_cast<const uint32_t *>(pb + P - 4));
    uint32_t sh_;
    ph = decode4(hw, sh_);
    // W: bit (18-i) <-> position P-3+i is a valid base of this contig below seg_hi
    uint32_t W = ~((bad4(sh_) & 7u) << 16 | bad4(s0) << 12 | bad4(s1) << 8 | bad4(s2) << 4 | bad4(s3)) & 0x7FFFFu;
    if (P < 4 && !head_ok) W &= 0xFFFFu;
    uint32_t own;
    tnf_range_masks(P, o0, seg_lo, seg_hi, W, own);
    return tnf_masked_prep(cw, W, own, K);
}

// The warp counts every group with K != 0 together: lane b takes the 4-mer ending at base 15-b, so a masked group
// costs one shared-memory reduction instead of 16 predicated ones.  Called by the whole warp.
__device__ __forceinline__ void tnf_count_masked(uint32_t cw, uint32_t ph, uint32_t K, uint32_t t4, int lane) {
    unsigned todo = __ballot_sync(0xffffffffu, K != 0);
    while (todo) {
        const int src = __ffs(todo) - 1;
        todo &= todo - 1;
        const uint32_t cb = __shfl_sync(0xffffffffu, cw, src), hb = __shfl_sync(0xffffffffu, ph, src);
        const uint32_t kb = __shfl_sync(0xffffffffu, K, src);
        if ((kb >> lane) & 1u) {  // K has 16 bits: lanes 16..31 never count
            const uint32_t t = lane ? __funnelshift_r(cb, hb, 2 * lane - 2) : (cb << 2);
            smem_inc(t4 + (t & 0x3FCu));
        }
    }
}

// Row with at least one group that is not clean (a non-ACGT byte).  Lanes whose group is clean take the vector path;
// the others compute the exact per-base validity of their group, and then the warp counts those groups together.
// Called by the whole warp, out of line so that the streaming loop stays small.  Returns the number of G/C bases
// counted for this lane.
template <int MODE>
__device__ __noinline__ uint32_t tnf_edge_row(uint4 v, uint32_t cw, uint32_t halo, bool fast, int g, int g_last,
                                              int o0, int seg_lo, int seg_hi, const uint8_t *__restrict__ pb,
                                              bool head_ok) {
    const uint32_t t5 = (uint32_t)__cvta_generic_to_shared(tnf_tab);
    const uint32_t t4 = t5 + 4u * TnfTab<MODE>::T5W;
    const int lane = threadIdx.x;
    uint32_t gcn = 0, K = 0, ph = 0;
    if (g <= g_last) {
        if (fast) {
            tnf_count16<MODE>(cw, halo, t5, t4);
            gcn = __popc((cw ^ (cw >> 1)) & 0x55555555u);
        } else {
            uint32_t s0, s1, s2, s3;
            decode4_hi(v.x, s0); decode4_hi(v.y, s1); decode4_hi(v.z, s2); decode4_hi(v.w, s3);
            gcn = tnf_group_masks(s0, s1, s2, s3, cw, g, o0, seg_lo, seg_hi, pb, head_ok, K, ph);
        }
    }
    tnf_count_masked(cw, ph, K, t4, lane);
    return gcn;
}

template <int MODE, int PS>
__global__ void __launch_bounds__(TNF_THREADS, 32)  // 32 warps per SM: at most 64 registers
tnf_kernel(const uint8_t *__restrict__ bases_al, int mis, int64_t end_q,
           const int64_t *__restrict__ offsets, int64_t n,
           const int32_t *__restrict__ first_contig, const int32_t *__restrict__ adj, int64_t chunk_begin,
           int64_t n_chunks, int64_t total_pieces,
           uint32_t *__restrict__ counts, unsigned long long *__restrict__ gc_counts,
           int *__restrict__ work_counter) {
    constexpr int WORDS = TnfTab<MODE>::WORDS;
    constexpr int64_t TNF_PIECE = 1 << PS;
    uint32_t *tab = tnf_tab;

    const int lane = threadIdx.x;
    // opaque to the compiler: a plain cvta result is rematerialised per row (S2UR SR_CgaCtaId + UMOV + ULEA)
    uint32_t t5 = (uint32_t)__cvta_generic_to_shared(tnf_tab);  // T5 (MODE 5) or T4 (MODE 4)
    asm volatile("mov.u32 %0, %0;" : "+r"(t5));
    const uint32_t t4 = t5 + 4u * TnfTab<MODE>::T5W;
    const int src_lane = (lane + 31) & 31;

    for (int i = lane; i < WORDS; i += 32) tab[i] = 0;
    uint32_t zero_phase = 0;  // parity of the bulk-clear barrier's current phase
    tnf_bar_init(lane);
    __syncwarp();

    for (;;) {
        int64_t kc = 0;
        if (lane == 0) kc = atomicAdd(work_counter, 1);
        kc = __shfl_sync(0xffffffffu, kc, 0);
        if (kc >= n_chunks) break;
        // the piece = aligned-stream positions [lo_abs, hi_abs): nominal boundaries moved to contig starts
        const int64_t kabs = chunk_begin + kc;
        int64_t lo_abs = kabs * TNF_PIECE + adj[kc];
        int64_t hi_abs = kabs + 1 >= total_pieces ? end_q : (kabs + 1) * TNF_PIECE + adj[kc + 1];
        lo_abs = lo_abs < mis ? mis : lo_abs;
        hi_abs = hi_abs < end_q ? hi_abs : end_q;
        const int64_t base = lo_abs & ~(int64_t)15;
        const uint8_t *__restrict__ pb = bases_al + base;     // 16-byte aligned
        // everything below is relative to `base`, in 32 bits
        const int piece_lo = (int)(lo_abs - base);
        const int piece_hi = (int)(hi_abs - base);
        // bytes at or beyond safe_hi may not be read: the group holding the piece's last byte is read whole
        // unless it runs past the end of the stream
        const int64_t room = end_q - base;
        const int safe_hi = (int)(room < ((piece_hi + 15) & ~15) ? room : ((piece_hi + 15) & ~15));
        const bool head_ok = base > 0; // bytes before the piece may be read (halo)

        // the contig boundaries are fetched one segment ahead (a segment's first loads wait for them otherwise)
        int64_t c = first_contig[kc];
        int64_t off_c = offsets[c], off_n = offsets[c + 1], off_nn = 0;
        for (; c < n; c++, off_c = off_n, off_n = off_nn) {
            off_nn = c + 2 <= n ? offsets[c + 2] : off_n;
            const int64_t o0_64 = off_c + mis - base;
            if (o0_64 >= piece_hi) break;
            const int64_t o1_64 = off_n + mis - base;
            const int o0 = o0_64 < -64 ? -64 : (int)o0_64;  // only compared with positions >= -3
            const int seg_lo = o0 > piece_lo ? o0 : piece_lo;
            const int seg_hi = o1_64 < piece_hi ? (int)o1_64 : piece_hi;
            if (seg_hi <= seg_lo) continue;

            uint32_t gcl = 0;
            const int g_first = seg_lo >> 4, g_last = (seg_hi - 1) >> 4;
            // interior groups: all 16 bases in the segment and the 3 halo bases in the contig
            const int in_lo = o0 + 3 > seg_lo ? o0 + 3 : seg_lo;
            const int g_in_lo = (in_lo + 15) >> 4;
            const int g_in_hi = (seg_hi >> 4) > g_in_lo ? (seg_hi >> 4) : g_in_lo;  // [g_in_lo, g_in_hi)
            // The rows cover the INTERIOR groups only (whole groups inside the segment: always safe to load); the at
            // most three partial groups of the segment (head: up to two, tail: one) are counted together afterwards.
            // So the first row of a segment is an ordinary row, the last one an ordinary row with idle lanes, and the
            // per-base masks are built once per segment instead of once per edge row.
            if (g_in_hi > g_in_lo) {
                // halo of a group = last 3 codes of the group before it | bit 6 "not usable", 7 bits.  The one before
                // g_in_lo is kept by lane 31; its bytes lie inside the contig (16 g_in_lo >= o0 + 3) and inside the
                // buffer (g_in_lo >= 1).
                uint32_t carry = 0x40u;
                TNF_CHECK(c >= 0 && c < n && base + ((int64_t)g_in_lo << 4) - 4 >= 0 &&
                          ((int64_t)g_in_hi << 4) <= room);
                if (lane == 31) {
                    uint32_t sg;
                    const uint32_t ph = decode4(__ldg(reinterpret_cast<const uint32_t *>(pb + (g_in_lo << 4) - 4)), sg);
                    carry = (ph & 0x3Fu) | (((sg ^ 0x41414141u) & 0xFFFFFF00u) ? 0x40u : 0u);
                }
                int g0 = g_in_lo;
                // ---- whole tiles: TNF_UNROLL rows decoded together, ONE shuffle for their halos (the shuffle runs on the
                // pipe that bounds the kernel: 4 % of its wavefronts were per-row shuffles) ----
                for (; g0 + 32 * TNF_UNROLL <= g_in_hi; g0 += 32 * TNF_UNROLL) {
                    uint32_t cw[TNF_UNROLL];
                    uint32_t pk = 0;  // this lane's TNF_UNROLL halos-to-give, 7 bits each
                    {
                        uint4 v[TNF_UNROLL];
#pragma unroll
                        for (int u = 0; u < TNF_UNROLL; u++) {
                            TNF_CHECK(g0 + 32 * u + lane >= 0 && (((int64_t)(g0 + 32 * u + lane) << 4) + 16) <= room);
                            v[u] = __ldg(reinterpret_cast<const uint4 *>(pb + ((g0 + 32 * u + lane) << 4)));
                        }
#pragma unroll
                        for (int u = 0; u < TNF_UNROLL; u++) {
                            uint32_t s0, s1, s2, s3;
                            const uint32_t m0 = decode4_hi(v[u].x, s0), m1 = decode4_hi(v[u].y, s1);
                            const uint32_t m2 = decode4_hi(v[u].z, s2), m3 = decode4_hi(v[u].w, s3);
                            cw[u] = __byte_perm(__byte_perm(m3, m2, 0x0073), __byte_perm(m1, m0, 0x0073), 0x5410);
                            const bool clean = (((s0 ^ s1) | (s2 ^ s3) | (s0 ^ s2)) == 0) && (s0 == 0x41414141u);
                            pk |= ((cw[u] & 0x3Fu) | (clean ? 0u : 0x40u)) << (7 * u);
                        }
                    }
                    // lane L takes the halos of its groups from lane L - 1; lane 0's come from lane 31's groups one row
                    // earlier: the carry and this tile's rows 0 .. TNF_UNROLL - 2
                    const uint32_t got = __shfl_sync(0xffffffffu, lane == 31 ? (carry | (pk << 7)) : pk, src_lane);
                    carry = (pk >> (7 * (TNF_UNROLL - 1))) & 0x7Fu;  // only lane 31's copy is ever consumed
                    // a group is fast when it is clean and so is the one before it: bit 6 of its own and of the received field
                    if (__all_sync(0xffffffffu, ((pk | got) & 0x08102040u) == 0)) {  // the whole tile: one vote
#pragma unroll
                        for (int u = 0; u < TNF_UNROLL; u++) {
                            gcl += __popc((cw[u] ^ (cw[u] >> 1)) & 0x55555555u);
                            tnf_count16<MODE>(cw[u], (got >> (7 * u)) & 0x3Fu, t5, t4);
                        }
                    } else {
#pragma unroll
                        for (int u = 0; u < TNF_UNROLL; u++) {
                            const uint32_t halo = (got >> (7 * u)) & 0x7Fu;
                            const bool fast = !((pk >> (7 * u)) & 0x40u) && halo < 0x40u;
                            if (__all_sync(0xffffffffu, fast)) {
                                gcl += __popc((cw[u] ^ (cw[u] >> 1)) & 0x55555555u);
                                tnf_count16<MODE>(cw[u], halo, t5, t4);
                            } else {  // a group with a non-ACGT byte: its bytes are fetched again (L1) for the exact masks
                                const int g = g0 + 32 * u + lane;
                                TNF_CHECK(g < g_in_hi);
                                const uint4 vv = __ldg(reinterpret_cast<const uint4 *>(pb + (g << 4)));
                                gcl += tnf_edge_row<MODE>(vv, cw[u], halo, fast, g, g_in_hi - 1, o0, seg_lo, seg_hi, pb, head_ok);
                            }
                        }
                    }
                }
                // ---- the segment's last, partial tile: row by row ----
                if (g0 < g_in_hi) {
                    uint4 v[TNF_UNROLL];
#pragma unroll
                    for (int u = 0; u < TNF_UNROLL; u++) {
                        const int g = g0 + 32 * u + lane;
                        v[u] = make_uint4(0, 0, 0, 0);
                        if (g < g_in_hi) v[u] = __ldg(reinterpret_cast<const uint4 *>(pb + (g << 4)));
                    }
#pragma unroll
                    for (int u = 0; u < TNF_UNROLL; u++) {
                        if (g0 + 32 * u >= g_in_hi) break;  // warp-uniform
                        const int g = g0 + 32 * u + lane;
                        uint32_t s0, s1, s2, s3;
                        const uint32_t m0 = decode4_hi(v[u].x, s0), m1 = decode4_hi(v[u].y, s1);
                        const uint32_t m2 = decode4_hi(v[u].z, s2), m3 = decode4_hi(v[u].w, s3);
                        const uint32_t cw = __byte_perm(__byte_perm(m3, m2, 0x0073), __byte_perm(m1, m0, 0x0073), 0x5410);
                        // 16 valid bases (a group beyond the interior was loaded as zeros and is never clean)
                        const bool clean = (((s0 ^ s1) | (s2 ^ s3) | (s0 ^ s2)) == 0) && (s0 == 0x41414141u);
                        const uint32_t mine = (cw & 0x3Fu) | (clean ? 0u : 0x40u);
                        const uint32_t halo = __shfl_sync(0xffffffffu, lane == 31 ? carry : mine, src_lane);
                        carry = mine;
                        const bool fast = clean && (halo < 0x40u);
                        if (__all_sync(0xffffffffu, fast)) {
                            gcl += __popc((cw ^ (cw >> 1)) & 0x55555555u);
                            tnf_count16<MODE>(cw, halo, t5, t4);
                        } else if (__all_sync(0xffffffffu, fast || g >= g_in_hi)) {  // the segment's last row
                            if (g < g_in_hi) {
                                gcl += __popc((cw ^ (cw >> 1)) & 0x55555555u);
                                tnf_count16<MODE>(cw, halo, t5, t4);
                            }
                        } else {
                            gcl += tnf_edge_row<MODE>(v[u], cw, halo, fast, g, g_in_hi - 1, o0, seg_lo, seg_hi, pb, head_ok);
                        }
                    }
                }
            }
            {   // partial groups: g_first .. min(g_in_lo, g_last + 1) - 1 and, if the segment ends inside a group, g_in_hi
                const int n_head = (g_in_lo < g_last + 1 ? g_in_lo : g_last + 1) - g_first;
                const bool has_tail = g_in_hi <= g_last;
                if (n_head > 0 || has_tail) {  // warp-uniform
                    const int g = lane < n_head ? g_first + lane : (lane == n_head && has_tail) ? g_in_hi : INT_MAX;
                    uint4 vv = make_uint4(0, 0, 0, 0);
                    if (g != INT_MAX) {
                        const int P = g << 4;
                        TNF_CHECK(g >= 0 && g <= g_last && P < safe_hi && (int64_t)safe_hi <= room);
                        vv = P + 16 <= safe_hi ? __ldg(reinterpret_cast<const uint4 *>(pb + P)) : tnf_load_tail(pb, P, safe_hi);
                    }
                    uint32_t s0, s1, s2, s3;
                    const uint32_t m0 = decode4_hi(vv.x, s0), m1 = decode4_hi(vv.y, s1);
                    const uint32_t m2 = decode4_hi(vv.z, s2), m3 = decode4_hi(vv.w, s3);
                    const uint32_t cw = __byte_perm(__byte_perm(m3, m2, 0x0073), __byte_perm(m1, m0, 0x0073), 0x5410);
                    uint32_t K = 0, ph = 0;
                    if (g != INT_MAX) gcl += tnf_group_masks(s0, s1, s2, s3, cw, g, o0, seg_lo, seg_hi, pb, head_ok, K, ph);
                    tnf_count_masked(cw, ph, K, t4, lane);
                }
            }
            gcl = __reduce_add_sync(0xffffffffu, gcl);
            const bool whole = o0_64 >= piece_lo && o1_64 <= piece_hi;  // the whole contig is this segment
            if (lane == 0) {
                if (whole) gc_counts[c] = (unsigned long long)gcl;
                else if (gcl) atomicAdd(&gc_counts[c], (unsigned long long)gcl);
            }
            tnf_flush<MODE>(tab, lane, counts + c * MP2_NCANON, zero_phase, whole);
        }
    }
}

// ---- 2-bit packed input (north_star layout) ---------------------------------------------------
// codes: 16 bases per uint32, FIRST base most significant; valid: 32 bases per uint32, first base
// most significant (0 = any non-ACGT byte).  Same piece / segment / table scheme as tnf_kernel; a
// lane owns QUADS (4 consecutive groups = 64 bases: one 128-bit load of codes + one 64-bit load of
// validity), so the halo of groups 1..3 is already in registers and one shuffle serves 64 bases.
template <int MODE, int PS>
__global__ void __launch_bounds__(TNF_THREADS)
tnf_packed_kernel(const uint32_t *__restrict__ codes, const uint32_t *__restrict__ valid, int64_t n_bases,
                  const int64_t *__restrict__ offsets, int64_t n, const int32_t *__restrict__ first_contig,
                  int64_t chunk_begin, int64_t n_chunks, uint32_t *__restrict__ counts,
                  unsigned long long *__restrict__ gc_counts, int *__restrict__ work_counter) {
    constexpr int WORDS = TnfTab<MODE>::WORDS;
    constexpr int TNF_PIECE = 1 << PS;
    uint32_t *tab = tnf_tab;
    const int lane = threadIdx.x;
    uint32_t t5 = (uint32_t)__cvta_generic_to_shared(tnf_tab);
    asm volatile("mov.u32 %0, %0;" : "+r"(t5));  // opaque: not rematerialised per use (see tnf_kernel)
    const uint32_t t4 = t5 + 4u * TnfTab<MODE>::T5W;
    const int src_lane = (lane + 31) & 31;
    for (int i = lane; i < WORDS; i += 32) tab[i] = 0;
    uint32_t zero_phase = 0;
    tnf_bar_init(lane);
    __syncwarp();

    for (;;) {
        int64_t kc = 0;
        if (lane == 0) kc = atomicAdd(work_counter, 1);
        kc = __shfl_sync(0xffffffffu, kc, 0);
        if (kc >= n_chunks) break;
        const int64_t base = (chunk_begin + kc) * TNF_PIECE;
        const uint32_t *__restrict__ pc = codes + (base >> 4);
        const uint32_t *__restrict__ pv = valid + (base >> 5);
        const int piece_hi = (int)(n_bases - base < TNF_PIECE ? n_bases - base : TNF_PIECE);

        for (int64_t c = first_contig[kc]; c < n; c++) {
            const int64_t o0_64 = offsets[c] - base;
            if (o0_64 >= piece_hi) break;
            const int64_t o1_64 = offsets[c + 1] - base;
            const int o0 = o0_64 < -64 ? -64 : (int)o0_64;
            const int seg_lo = o0 > 0 ? o0 : 0;
            const int seg_hi = o1_64 < piece_hi ? (int)o1_64 : piece_hi;
            if (seg_hi <= seg_lo) continue;

            uint32_t gcl = 0;
            const int g_first = seg_lo >> 4, g_last = (seg_hi - 1) >> 4;
            const int in_lo = o0 + 3 > seg_lo ? o0 + 3 : seg_lo;
            const int g_in_lo = (in_lo + 15) >> 4;
            const int g_in_hi = (seg_hi >> 4) > g_in_lo ? (seg_hi >> 4) : g_in_lo;
            const int q_first = g_first >> 2, q_last = g_last >> 2;
            // packed halo = 6 code bits | 3 validity bits << 6 of the group before the first quad
            uint32_t carry = 0;
            if (lane == 31 && (q_first > 0 || base > 0)) {
                const int gp = 4 * q_first - 1;
                const uint32_t cwp = __ldg(pc + gp);
                const uint32_t vwp = __ldg(pv + (gp >> 1));  // arithmetic shift: gp = -1 -> word -1
                const uint32_t v16 = (gp & 1) ? (vwp & 0xFFFFu) : (vwp >> 16);
                carry = (cwp & 0x3Fu) | ((v16 & 7u) << 6);
            }
            for (int q0 = q_first; q0 <= q_last; q0 += 32) {
                const int qd = q0 + lane;
                uint4 cq = make_uint4(0, 0, 0, 0);
                uint2 vq = make_uint2(0, 0);
                if (qd <= q_last) {
                    cq = __ldg(reinterpret_cast<const uint4 *>(pc) + qd);
                    vq = __ldg(reinterpret_cast<const uint2 *>(pv) + qd);
                }
                const uint32_t cws[4] = {cq.x, cq.y, cq.z, cq.w};
                const uint32_t v16s[4] = {vq.x >> 16, vq.x & 0xFFFFu, vq.y >> 16, vq.y & 0xFFFFu};
                const uint32_t mine = (cws[3] & 0x3Fu) | ((v16s[3] & 7u) << 6);
                const uint32_t give = lane == 31 ? carry : mine;
                uint32_t halo = __shfl_sync(0xffffffffu, give, src_lane);
                carry = mine;
                if (qd > q_last) continue;
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    const int g = 4 * qd + k;
                    const uint32_t cw = cws[k], v16 = v16s[k];
                    if (g >= g_first && g <= g_last) {
                        const bool fast = (v16 == 0xFFFFu) && ((halo >> 6) == 7u) &&
                                          ((unsigned)(g - g_in_lo) < (unsigned)(g_in_hi - g_in_lo));
                        if (fast) {
                            gcl += __popc((cw ^ (cw >> 1)) & 0x55555555u);
                            tnf_count16<MODE>(cw, halo & 0x3Fu, t5, t4);
                        } else {
                            const int P = g << 4;
                            uint32_t W = ((halo >> 6) << 16) | v16, own;
                            tnf_range_masks(P, o0, seg_lo, seg_hi, W, own);
                            gcl += tnf_masked16(cw, halo & 0x3Fu, W, own, t4);
                        }
                    }
                    halo = (cw & 0x3Fu) | ((v16 & 7u) << 6);
                }
            }
            gcl = __reduce_add_sync(0xffffffffu, gcl);
            if (lane == 0 && gcl) atomicAdd(&gc_counts[c], (unsigned long long)gcl);
            tnf_flush<MODE>(tab, lane, counts + c * MP2_NCANON, zero_phase, false);
        }
    }
}

// ASCII -> (codes, valid).  One thread per group of 16 bases; groups are padded to whole quads.
__global__ void tnf_pack_kernel(const uint8_t *__restrict__ bases, int64_t n_bases, int64_t n_groups,
                                uint32_t *__restrict__ codes, uint16_t *__restrict__ valid16) {
    const int64_t g = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (g >= n_groups) return;
    const int64_t P = g << 4;
    uint32_t w[4] = {0, 0, 0, 0};
    if (P + 16 <= n_bases && ((reinterpret_cast<uintptr_t>(bases + P) & 15) == 0)) {
        const uint4 v = __ldg(reinterpret_cast<const uint4 *>(bases + P));
        w[0] = v.x; w[1] = v.y; w[2] = v.z; w[3] = v.w;
    } else {
        for (int i = 0; i < 16 && P + i < n_bases; i++) w[i >> 2] |= (uint32_t)bases[P + i] << (8 * (i & 3));
    }
    uint32_t s0, s1, s2, s3;
    const uint32_t m0 = decode4_hi(w[0], s0), m1 = decode4_hi(w[1], s1);
    const uint32_t m2 = decode4_hi(w[2], s2), m3 = decode4_hi(w[3], s3);
    const uint32_t cw = __byte_perm(__byte_perm(m3, m2, 0x0073), __byte_perm(m1, m0, 0x0073), 0x5410);
    uint32_t v16 = ~(bad4(s0) << 12 | bad4(s1) << 8 | bad4(s2) << 4 | bad4(s3)) & 0xFFFFu;
    if (P + 16 > n_bases) {
        const int keep = n_bases > P ? (int)(n_bases - P) : 0;  // first `keep` bases exist
        v16 &= keep ? (0xFFFFu << (16 - keep)) & 0xFFFFu : 0u;
    }
    // the code bits of an invalid base are cleared so the packed form is canonical
    uint32_t sp = v16;
    sp = (sp | (sp << 8)) & 0x00FF00FFu;
    sp = (sp | (sp << 4)) & 0x0F0F0F0Fu;
    sp = (sp | (sp << 2)) & 0x33333333u;
    sp = (sp | (sp << 1)) & 0x55555555u;
    codes[g] = cw & (sp | (sp << 1));
    valid16[g ^ 1] = (uint16_t)v16;  // even group = upper half of the little-endian 32-bit word
}

// ---- finalisation --------------------------------------------------------------------------
// relative frequencies (src/composition.rs:111-118): u32 -> f32, row sum in f32, divide,
// NaN -> 0.  The f32 row sum equals the exact integer sum while it is < 2^24; above that the
// lane-0 fallback reproduces the 8-way unrolled fold order of ndarray 0.13 (numeric_util::unrolled_fold).
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

// MP2_TNF_ZERO_ROWS=0: mp2_tnf_device clears all rows with a memset instead of the few that need it (A-B switch).
static bool tnf_zero_rows() {
    static int v = -1;
    if (v < 0) {
        const char *e = getenv("MP2_TNF_ZERO_ROWS");
        v = (e && e[0] == '0') ? 0 : 1;
    }
    return v == 1;
}

// MP2_TNF_SNAP=0 keeps the nominal piece boundaries (development / A-B switch).
static int tnf_snap() {
    static int snap = -1;
    if (snap < 0) {
        const char *e = getenv("MP2_TNF_SNAP");
        snap = (e && e[0] == '0') ? 0 : 1;
    }
    return snap;
}

int64_t tnf_num_chunks(const void *d_bases, int64_t n_bases);

// Rows the counting kernel does not store whole: contigs that a piece boundary cuts (their segments add into the row)
// and empty contigs (no segment at all).  Every other row is written by the one flush of its contig, so a launch over
// the whole stream clears these few rows instead of all of them (163 MB and 50 us of a 1.4 ms step at C3).
// One warp per boundary k = 1 .. n_chunks - 1 (adj / first_contig as tnf_chunk_map_kernel left them, chunk_begin = 0).
__global__ void tnf_zero_split_rows_kernel(const int64_t *__restrict__ offsets, int mis, int piece_shift, int64_t n_chunks,
                                           const int32_t *__restrict__ first_contig, const int32_t *__restrict__ adj,
                                           uint32_t *__restrict__ counts, unsigned long long *__restrict__ gc_counts) {
    const int64_t k = ((blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5) + 1;
    const int lane = threadIdx.x & 31;
    if (k >= n_chunks) return;
    const int64_t pos = (k << piece_shift) - mis + adj[k];  // stream position of the boundary
    const int64_t c = first_contig[k];                      // largest c with offsets[c] <= pos
    if (offsets[c] < pos && pos < offsets[c + 1]) {
        for (int j = lane; j < MP2_NCANON; j += 32) counts[c * MP2_NCANON + j] = 0;
        if (lane == 0) gc_counts[c] = 0;
    }
}

__global__ void tnf_zero_empty_rows_kernel(const int64_t *__restrict__ offsets, int64_t n, uint32_t *__restrict__ counts,
                                           unsigned long long *__restrict__ gc_counts) {
    const int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (c >= n || offsets[c + 1] > offsets[c]) return;
    for (int j = 0; j < MP2_NCANON; j++) counts[c * MP2_NCANON + j] = 0;
    gc_counts[c] = 0;
}

// Enqueue the counting kernels for chunks [chunk_begin, chunk_end) of the stream.
// d_counts / d_gc_counts must have been zeroed (once) before the first range -- or, for ONE launch over the whole
// stream, `zero_rows` = true clears just the rows that need it (see above).
int tnf_launch_range(const uint8_t *d_bases, const int64_t *d_offsets, int64_t n, int64_t n_bases,
                     int64_t chunk_begin, int64_t chunk_end, uint32_t *d_counts,
                     unsigned long long *d_gc_counts, cudaStream_t stream, bool zero_rows) {
    MP2_TRY(upload_tables());
    const int mis = (int)(reinterpret_cast<uintptr_t>(d_bases) & 15);
    const int64_t n_chunks = chunk_end - chunk_begin;
    if (n_chunks <= 0) return MP2_OK;
    const int ps = tnf_piece_shift(n_bases);
    Temp map, adj, ctr;
    MP2_TRY(map.alloc(sizeof(int32_t) * (n_chunks + 1), stream));
    MP2_TRY(adj.alloc(sizeof(int32_t) * (n_chunks + 1), stream));
    MP2_TRY(ctr.alloc(sizeof(int), stream));
    MP2_CUDA(cudaMemsetAsync(ctr.p, 0, sizeof(int), stream));
    const int64_t total_pieces = tnf_num_chunks(d_bases, n_bases);
    {
        Launch L("tnf_chunk_map_kernel", stream);
        tnf_chunk_map_kernel<<<(unsigned)((n_chunks + 1 + 255) / 256), 256, 0, stream>>>(
            d_offsets, n, mis, ps, tnf_snap(), chunk_begin, n_chunks, map.as<int32_t>(), adj.as<int32_t>());
    }
    MP2_CUDA(cudaGetLastError());
    if (zero_rows) {
        if (chunk_begin != 0 || chunk_end != total_pieces) return fail(MP2_EINVAL, "zero_rows needs a launch over the whole stream");
        Launch L("tnf_zero_rows_kernel", stream);
        if (n_chunks > 1)
            tnf_zero_split_rows_kernel<<<(unsigned)(((n_chunks - 1) * 32 + 255) / 256), 256, 0, stream>>>(
                d_offsets, mis, ps, n_chunks, map.as<int32_t>(), adj.as<int32_t>(), d_counts, d_gc_counts);
        tnf_zero_empty_rows_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(d_offsets, n, d_counts, d_gc_counts);
    }
    MP2_CUDA(cudaGetLastError());
    const int mode = tnf_mode();
#define MP2_TNF_LAUNCH(M, PS)                                                                               \
    do {                                                                                                    \
        int occ = 0;                                                                                        \
        MP2_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, tnf_kernel<M, PS>, TNF_THREADS, 0));   \
        if (occ < 1) occ = 1;                                                                               \
        int64_t grid = (int64_t)num_sms() * occ;                                                            \
        if (grid > n_chunks) grid = n_chunks;                                                               \
        Launch L("tnf_kernel", stream);                                                                     \
        tnf_kernel<M, PS><<<(unsigned)grid, TNF_THREADS, 0, stream>>>(                                      \
            d_bases - mis, mis, n_bases + mis, d_offsets, n, map.as<int32_t>(), adj.as<int32_t>(),         \
            chunk_begin, n_chunks, total_pieces, d_counts, d_gc_counts, ctr.as<int>());                     \
    } while (0)
    if (mode == 5 && ps == 16) MP2_TNF_LAUNCH(5, 16);
    else if (mode == 5 && ps == 15) MP2_TNF_LAUNCH(5, 15);
    else if (mode == 5) MP2_TNF_LAUNCH(5, 14);
    else if (ps == 16) MP2_TNF_LAUNCH(4, 16);
    else if (ps == 15) MP2_TNF_LAUNCH(4, 15);
    else MP2_TNF_LAUNCH(4, 14);
#undef MP2_TNF_LAUNCH
    MP2_CUDA(cudaGetLastError());
    return MP2_OK;
}

// Pieces wholly inside the first `done` bytes of the stream (host staging pipelines on this).
int64_t tnf_chunks_ready(const void *d_bases, int64_t done, int64_t n_bases) {
    // a piece may reach one piece beyond its nominal end (boundaries snap to contig starts) and its last
    // 16-byte group is read whole
    const int mis = (int)(reinterpret_cast<uintptr_t>(d_bases) & 15);
    const int ps = tnf_piece_shift(n_bases);
    const int64_t usable = done + mis - ((int64_t)1 << ps) - 16;
    return usable > 0 ? usable >> ps : 0;
}

int64_t tnf_num_chunks(const void *d_bases, int64_t n_bases) {
    const int mis = (int)(reinterpret_cast<uintptr_t>(d_bases) & 15);
    const int ps = tnf_piece_shift(n_bases);
    return (n_bases + mis + ((int64_t)1 << ps) - 1) >> ps;
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
    if (n_bases == 0 || !tnf_zero_rows()) {  // (no piece, no flush: every row is empty)
        MP2_CUDA(cudaMemsetAsync(d_counts, 0, sizeof(uint32_t) * MP2_NCANON * n, stream));
        MP2_CUDA(cudaMemsetAsync(gcc.p, 0, sizeof(unsigned long long) * n, stream));
    }
    MP2_TRY(tnf_launch_range((const uint8_t *)d_bases, (const int64_t *)d_offsets, n, n_bases, 0,
                             tnf_num_chunks(d_bases, n_bases), (uint32_t *)d_counts,
                             gcc.as<unsigned long long>(), stream, tnf_zero_rows()));
    MP2_TRY(tnf_finalize((const uint32_t *)d_counts, gcc.as<unsigned long long>(),
                         (const int64_t *)d_offsets, n, (float *)d_rel, (float *)d_gc, stream));
    return MP2_OK;
}

extern "C" int mp2_tnf_normalize_device(const void *d_counts, int64_t n, void *d_rel, void *stream_) {
    using namespace mp2;
    if (n < 0) return fail(MP2_EINVAL, "negative size");
    if (!d_counts || !d_rel) return fail(MP2_EINVAL, "NULL argument");
    MP2_TRY(require_device(-1));
    return tnf_finalize((const uint32_t *)d_counts, nullptr, nullptr, n, (float *)d_rel, nullptr, (cudaStream_t)stream_);
}

extern "C" int mp2_pack_device(const void *d_bases, int64_t n_bases, void *d_codes, void *d_valid, void *stream_) {
    using namespace mp2;
    if (n_bases < 0) return fail(MP2_EINVAL, "negative size");
    if (!d_codes || !d_valid || (n_bases > 0 && !d_bases)) return fail(MP2_EINVAL, "NULL argument");
    if ((reinterpret_cast<uintptr_t>(d_codes) & 15) || (reinterpret_cast<uintptr_t>(d_valid) & 7))
        return fail(MP2_EINVAL, "d_codes must be 16-byte and d_valid 8-byte aligned");
    MP2_TRY(require_device(-1));
    cudaStream_t stream = (cudaStream_t)stream_;
    const int64_t n_groups = ((n_bases + 63) / 64) * 4;
    if (n_groups == 0) return MP2_OK;
    {
        Launch L("tnf_pack_kernel", stream);
        tnf_pack_kernel<<<(unsigned)((n_groups + 255) / 256), 256, 0, stream>>>(
            (const uint8_t *)d_bases, n_bases, n_groups, (uint32_t *)d_codes, (uint16_t *)d_valid);
    }
    MP2_CUDA(cudaGetLastError());
    return MP2_OK;
}

namespace mp2 {

// valid mask from the runs of invalid bases (the mask was memset to all ones): one warp per run, interior words
// are stored, the two edge words are and-ed.  First base of a word = most significant bit.
__global__ void tnf_valid_runs_kernel(const int64_t *__restrict__ runs, int64_t n_runs, uint32_t *__restrict__ valid) {
    const int lane = threadIdx.x & 31;
    const int64_t r = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    if (r >= n_runs) return;
    const int64_t b = runs[2 * r], e = runs[2 * r + 1];
    if (e <= b) return;
    const int64_t w0 = b >> 5, w1 = (e - 1) >> 5;
    for (int64_t w = w0 + lane; w <= w1; w += 32) {
        const int lo = w == w0 ? (int)(b & 31) : 0, hi = w == w1 ? (int)((e - 1) & 31) + 1 : 32;  // bases [lo, hi) of word w
        const uint32_t m = (hi - lo == 32) ? 0xFFFFFFFFu : (((1u << (hi - lo)) - 1u) << (32 - hi));
        if (m == 0xFFFFFFFFu) valid[w] = 0u;
        else atomicAnd(&valid[w], ~m);
    }
}

// d_valid (2 words per quad of 64 bases) from `n_runs` (begin, end) pairs on the device; positions from n_bases to
// the end of the last quad are invalid.
int tnf_valid_from_runs(const int64_t *d_runs, int64_t n_runs, int64_t n_bases, uint32_t *d_valid, cudaStream_t stream) {
    const int64_t nq = (n_bases + 63) / 64;
    if (nq == 0) return MP2_OK;
    MP2_CUDA(cudaMemsetAsync(d_valid, 0xFF, sizeof(uint32_t) * 2 * (size_t)nq, stream));
    if (n_runs > 0) {
        Launch L("tnf_valid_runs_kernel", stream);
        tnf_valid_runs_kernel<<<(unsigned)((n_runs * 32 + 255) / 256), 256, 0, stream>>>(d_runs, n_runs, d_valid);
    }
    MP2_CUDA(cudaGetLastError());
    return MP2_OK;
}

int64_t tnf_packed_num_chunks(int64_t n_bases) {
    const int ps = tnf_piece_shift(n_bases);
    return (n_bases + ((int64_t)1 << ps) - 1) >> ps;
}

// Bases covered by the first `chunks` pieces: every k-mer ending before that position has been counted.
int64_t tnf_packed_bases_done(int64_t chunks, int64_t n_bases) {
    const int64_t b = chunks << tnf_piece_shift(n_bases);
    return b < n_bases ? b : n_bases;
}

// Pieces wholly inside the first `done` bases (host staging pipelines on this): a piece reads its own quads and
// the last group of the quad before it.
int64_t tnf_packed_chunks_ready(int64_t done, int64_t n_bases) {
    return done >> tnf_piece_shift(n_bases);
}

// Enqueue the packed counting kernel for pieces [chunk_begin, chunk_end).  d_counts / d_gc_counts must have been
// zeroed (once) before the first range.
int tnf_packed_launch_range(const uint32_t *d_codes, const uint32_t *d_valid, const int64_t *d_offsets, int64_t n,
                            int64_t n_bases, int64_t chunk_begin, int64_t chunk_end, uint32_t *d_counts,
                            unsigned long long *d_gc_counts, cudaStream_t stream) {
    MP2_TRY(upload_tables());
    const int64_t n_chunks = chunk_end - chunk_begin;
    if (n_chunks <= 0) return MP2_OK;
    const int ps = tnf_piece_shift(n_bases);
    Temp map, adj, ctr;  // the packed kernel keeps the nominal boundaries (snap = 0)
    MP2_TRY(map.alloc(sizeof(int32_t) * (n_chunks + 1), stream));
    MP2_TRY(adj.alloc(sizeof(int32_t) * (n_chunks + 1), stream));
    MP2_TRY(ctr.alloc(sizeof(int), stream));
    MP2_CUDA(cudaMemsetAsync(ctr.p, 0, sizeof(int), stream));
    {
        Launch L("tnf_chunk_map_kernel", stream);
        tnf_chunk_map_kernel<<<(unsigned)((n_chunks + 1 + 255) / 256), 256, 0, stream>>>(
            d_offsets, n, 0, ps, 0, chunk_begin, n_chunks, map.as<int32_t>(), adj.as<int32_t>());
    }
    MP2_CUDA(cudaGetLastError());
    const int mode = tnf_mode();
#define MP2_TNFP_LAUNCH(M, PS)                                                                                  \
    do {                                                                                                        \
        int occ = 0;                                                                                            \
        MP2_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, tnf_packed_kernel<M, PS>, TNF_THREADS, 0));\
        if (occ < 1) occ = 1;                                                                                   \
        int64_t grid = (int64_t)num_sms() * occ;                                                                \
        if (grid > n_chunks) grid = n_chunks;                                                                   \
        Launch L("tnf_packed_kernel", stream);                                                                  \
        tnf_packed_kernel<M, PS><<<(unsigned)grid, TNF_THREADS, 0, stream>>>(                                   \
            d_codes, d_valid, n_bases, d_offsets, n, map.as<int32_t>(), chunk_begin, n_chunks, d_counts,        \
            d_gc_counts, ctr.as<int>());                                                                        \
    } while (0)
    if (mode == 5 && ps == 16) MP2_TNFP_LAUNCH(5, 16);
    else if (mode == 5 && ps == 15) MP2_TNFP_LAUNCH(5, 15);
    else if (mode == 5) MP2_TNFP_LAUNCH(5, 14);
    else if (ps == 16) MP2_TNFP_LAUNCH(4, 16);
    else if (ps == 15) MP2_TNFP_LAUNCH(4, 15);
    else MP2_TNFP_LAUNCH(4, 14);
#undef MP2_TNFP_LAUNCH
    MP2_CUDA(cudaGetLastError());
    return MP2_OK;
}

}  // namespace mp2

extern "C" int mp2_tnf_packed_device(const void *d_codes, const void *d_valid, const void *d_offsets, int64_t n,
                                     int64_t n_bases, void *d_counts, void *d_rel, void *d_gc, void *stream_) {
    using namespace mp2;
    if (n < 0 || n_bases < 0) return fail(MP2_EINVAL, "negative size");
    if (n > INT32_MAX) return fail(MP2_EINVAL, "too many sequences in one call (%lld)", (long long)n);
    if (!d_offsets || !d_counts || (n_bases > 0 && (!d_codes || !d_valid)))
        return fail(MP2_EINVAL, "d_codes, d_valid, d_offsets and d_counts are required");
    if ((reinterpret_cast<uintptr_t>(d_codes) & 15) || (reinterpret_cast<uintptr_t>(d_valid) & 7))
        return fail(MP2_EINVAL, "d_codes must be 16-byte and d_valid 8-byte aligned");
    MP2_TRY(require_device(-1));
    cudaStream_t stream = (cudaStream_t)stream_;
    if (n == 0) return MP2_OK;
    Temp gcc;
    MP2_TRY(gcc.alloc(sizeof(unsigned long long) * n, stream));
    MP2_CUDA(cudaMemsetAsync(d_counts, 0, sizeof(uint32_t) * MP2_NCANON * n, stream));
    MP2_CUDA(cudaMemsetAsync(gcc.p, 0, sizeof(unsigned long long) * n, stream));
    MP2_TRY(tnf_packed_launch_range((const uint32_t *)d_codes, (const uint32_t *)d_valid, (const int64_t *)d_offsets, n,
                                    n_bases, 0, tnf_packed_num_chunks(n_bases), (uint32_t *)d_counts,
                                    gcc.as<unsigned long long>(), stream));
    MP2_TRY(tnf_finalize((const uint32_t *)d_counts, gcc.as<unsigned long long>(), (const int64_t *)d_offsets, n,
                         (float *)d_rel, (float *)d_gc, stream));
    return MP2_OK;
}
