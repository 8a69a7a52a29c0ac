// Host-side 2-bit packer: ASCII bases -> (codes, valid), the north-star layout mp2_tnf_packed_device takes.
//
//   codes: 2 bits per base, 16 bases per uint32, FIRST base in the MOST significant bits (A=0 C=1 G=2 T=3,
//          case folded; the bits of a non-ACGT byte are 0)
//   valid: 1 bit per base, 32 bases per uint32, first base most significant (0 = N / IUPAC / anything else)
//
// Both arrays are padded to whole quads of 64 bases (4 code words + 2 validity words); positions beyond the
// end are invalid.  The input is the VIRTUAL concatenation of several buffers (the genomes of a run as the
// FASTA reader left them), so the run's stream is never materialised as ASCII: 0.375 bytes per base reach the
// pinned staging buffer instead of 1.  Replaces the sequence hand-over of magpurify2/core.py:35-50 +
// magpurify2/modules/composition.py:63-66 (Python str per contig) for the packed upload path.
//
// One pass, parallel over output quads (no two threads write the same word).  64 bytes per step with
// AVX-512BW when the CPU has it (two multiply-adds fold 4 bases into a byte), scalar otherwise; quads that
// straddle two input buffers or the end of the stream take the scalar path.
#include <immintrin.h>
#include <omp.h>
#include <stdint.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "mp2.h"

namespace {

inline void pack_quad_scalar(const uint8_t *b, int n, uint32_t *codes4, uint32_t *valid2) {
    uint32_t cw[4] = {0, 0, 0, 0}, vw[2] = {0, 0};
    for (int i = 0; i < n; i++) {
        const uint8_t x = b[i] & 0xDFu;  // case fold
        const bool ok = x == 'A' || x == 'C' || x == 'G' || x == 'T';
        if (ok) {
            const uint32_t code = ((b[i] >> 1) ^ (b[i] >> 2)) & 3u;
            cw[i >> 4] |= code << (2 * (15 - (i & 15)));
            vw[i >> 5] |= 1u << (31 - (i & 31));
        }
    }
    memcpy(codes4, cw, sizeof cw);
    memcpy(valid2, vw, sizeof vw);
}

inline uint32_t bitrev32(uint32_t v) {
    v = ((v >> 1) & 0x55555555u) | ((v & 0x55555555u) << 1);
    v = ((v >> 2) & 0x33333333u) | ((v & 0x33333333u) << 2);
    v = ((v >> 4) & 0x0F0F0F0Fu) | ((v & 0x0F0F0F0Fu) << 4);
    return __builtin_bswap32(v);
}

__attribute__((target("avx512f,avx512bw,avx512vl"))) inline void pack_quad_avx512(const uint8_t *b, uint32_t *codes4,
                                                                                  uint32_t *valid2) {
    const __m512i v = _mm512_loadu_si512((const void *)b);
    const __m512i up = _mm512_and_si512(v, _mm512_set1_epi8((char)0xDF));
    const __mmask64 ok = _mm512_cmpeq_epi8_mask(up, _mm512_set1_epi8('A')) | _mm512_cmpeq_epi8_mask(up, _mm512_set1_epi8('C')) |
                         _mm512_cmpeq_epi8_mask(up, _mm512_set1_epi8('G')) | _mm512_cmpeq_epi8_mask(up, _mm512_set1_epi8('T'));
    // code = ((b >> 1) ^ (b >> 2)) & 3 per byte (16-bit shifts: bits that cross a byte are masked away)
    const __m512i s1 = _mm512_srli_epi16(v, 1), s2 = _mm512_srli_epi16(v, 2);
    const __m512i c = _mm512_maskz_mov_epi8(ok, _mm512_and_si512(_mm512_xor_si512(s1, s2), _mm512_set1_epi8(3)));
    // 2 bases -> 4 bits (first base * 4 + second), 4 bases -> 8 bits
    const __m512i p2 = _mm512_maddubs_epi16(c, _mm512_set1_epi16(0x0104));
    const __m512i p4 = _mm512_madd_epi16(p2, _mm512_set1_epi32(0x00010010));
    // 16 dwords holding one byte each -> 16 bytes; byte-swap inside every 4 so that the first bases are on top
    const __m128i by = _mm512_cvtepi32_epi8(p4);
    const __m128i sw = _mm_shuffle_epi8(by, _mm_set_epi8(12, 13, 14, 15, 8, 9, 10, 11, 4, 5, 6, 7, 0, 1, 2, 3));
    _mm_storeu_si128((__m128i *)codes4, sw);
    valid2[0] = bitrev32((uint32_t)ok);
    valid2[1] = bitrev32((uint32_t)(ok >> 32));
}

}  // namespace

extern "C" int64_t mp2_packed_words(int64_t n_bases, int64_t *valid_words) {
    const int64_t quads = (n_bases + 63) / 64;
    if (valid_words) *valid_words = 2 * quads;
    return 4 * quads;
}

extern "C" int mp2_pack_host(const uint8_t *const *parts, const int64_t *part_len, int64_t n_parts, uint32_t *codes,
                             uint32_t *valid, int threads) {
    if (n_parts < 0 || (n_parts > 0 && (!parts || !part_len)) || !codes || !valid) return MP2_EINVAL;
    // start of every part in the virtual stream
    int64_t total = 0;
    for (int64_t p = 0; p < n_parts; p++) {
        if (part_len[p] < 0 || (part_len[p] > 0 && !parts[p])) return MP2_EINVAL;
        total += part_len[p];
    }
    const int64_t quads = (total + 63) / 64;
    if (quads == 0) return MP2_OK;
    if (threads < 1) threads = 1;
    const bool wide = __builtin_cpu_supports("avx512bw") && __builtin_cpu_supports("avx512vl");
    const int64_t grain = 4096;  // quads per task (256 KiB of bases)
    const int64_t tasks = (quads + grain - 1) / grain;
#pragma omp parallel for schedule(dynamic, 1) num_threads(threads)
    for (int64_t t = 0; t < tasks; t++) {
        const int64_t q0 = t * grain, q1 = std::min(quads, q0 + grain);
        // part holding the first base of this task (linear scan from the front is fine: few thousand parts at most
        // per task boundary; do it by binary search over an implicit prefix instead when there are many)
        int64_t p = 0, pstart = 0;
        {
            const int64_t pos = q0 * 64;
            while (p < n_parts && pstart + part_len[p] <= pos) pstart += part_len[p++];
        }
        for (int64_t q = q0; q < q1; q++) {
            const int64_t pos = q * 64;
            while (p < n_parts && pstart + part_len[p] <= pos) pstart += part_len[p++];
            uint32_t *cw = codes + 4 * q, *vw = valid + 2 * q;
            if (p < n_parts && pos + 64 <= pstart + part_len[p]) {  // the whole quad lies in one part
                const uint8_t *b = parts[p] + (pos - pstart);
                if (wide) pack_quad_avx512(b, cw, vw);
                else pack_quad_scalar(b, 64, cw, vw);
            } else {  // straddles parts, or the tail of the stream: gather the bytes first
                uint8_t tmp[64];
                int n = 0;
                int64_t pp = p, ps = pstart;
                while (n < 64 && pp < n_parts) {
                    const int64_t off = pos + n - ps;
                    const int64_t take = std::min<int64_t>(64 - n, part_len[pp] - off);
                    if (take > 0) {
                        memcpy(tmp + n, parts[pp] + off, (size_t)take);
                        n += (int)take;
                    }
                    if (n < 64) ps += part_len[pp++];
                }
                pack_quad_scalar(tmp, n, cw, vw);
            }
        }
    }
    return MP2_OK;
}

// Runs of invalid bases of a validity mask, as (begin, end) pairs in ascending order.  A genome is ACGT almost
// everywhere, so the runs are a few thousand pairs where the mask is 0.125 bytes per base: the upload path sends
// the codes and these pairs (0.25 bytes per base in all) and the device rebuilds the mask.  Bits beyond n_bases
// do not count.  Returns the number of runs, or -(needed) when `cap` pairs are not enough (nothing is written
// beyond cap), or MP2_EINVAL.
extern "C" int64_t mp2_invalid_runs(const uint32_t *valid, int64_t n_bases, int64_t *runs, int64_t cap, int threads) {
    if (n_bases < 0 || (n_bases > 0 && !valid) || cap < 0 || (cap > 0 && !runs)) return MP2_EINVAL;
    const int64_t words = (n_bases + 31) / 32;
    if (words == 0) return 0;
    if (threads < 1) threads = 1;
    const int64_t grain = 1 << 16;  // words per task (2 Mbases)
    const int64_t tasks = (words + grain - 1) / grain;
    std::vector<std::vector<int64_t>> local((size_t)tasks);
#pragma omp parallel for schedule(dynamic, 1) num_threads(threads)
    for (int64_t t = 0; t < tasks; t++) {
        std::vector<int64_t> &out = local[(size_t)t];
        const int64_t w0 = t * grain, w1 = std::min(words, w0 + grain);
        int64_t open = -1;  // begin of the run being extended
        for (int64_t w = w0; w < w1; w++) {
            uint32_t v = valid[w];
            const int64_t base = w * 32;
            if (base + 32 > n_bases) v |= (n_bases - base >= 32) ? 0u : (0xFFFFFFFFu >> (n_bases - base));  // beyond the end: not a run
            if (v == 0xFFFFFFFFu) {
                if (open >= 0) { out.push_back(open); out.push_back(base); open = -1; }
                continue;
            }
            uint32_t inv = ~v;  // first base = most significant bit
            int pos = 0;
            while (pos < 32) {
                const uint32_t rest = inv << pos;  // bit 31 = base `pos`
                if (open < 0) {
                    if (rest == 0) break;
                    const int z = __builtin_clz(rest);  // valid bases before the next invalid one
                    pos += z;
                    open = base + pos;
                } else {
                    const uint32_t ones = ~rest;  // bit 31 set <=> base `pos` valid
                    if (ones == 0) { pos = 32; break; }  // the run continues through the word
                    const int z = __builtin_clz(ones);     // invalid bases from pos on
                    pos += z;
                    if (pos < 32) { out.push_back(open); out.push_back(base + pos); open = -1; }
                }
            }
        }
        if (open >= 0) { out.push_back(open); out.push_back(std::min<int64_t>(w1 * 32, n_bases)); }
    }
    // concatenate, merging runs that touch across task (and word) boundaries
    int64_t n = 0, last_end = -1;
    for (int64_t t = 0; t < tasks; t++) {
        const std::vector<int64_t> &v = local[(size_t)t];
        for (size_t i = 0; i + 1 < v.size(); i += 2) {
            if (n > 0 && v[i] == last_end) {
                if (n <= cap) runs[2 * (n - 1) + 1] = v[i + 1];
            } else {
                if (n < cap) { runs[2 * n] = v[i]; runs[2 * n + 1] = v[i + 1]; }
                n++;
            }
            last_end = v[i + 1];
        }
    }
    return n <= cap ? n : -n;
}
