// Host-side BAM ingest: BGZF inflate (zlib, parallel over blocks) + record filter + alignment-block
// extraction.  Replaces htslib + coverm's ReferenceSortedBamFilter / CIGAR walk under
// _coverage.get_bam_coverages (/root/reference/src/coverage.rs:102-142).
// BAM decoding stays on host threads (north_star); the output (start, end) arrays are pinned so
// they can be copied to the device asynchronously.
#include <fcntl.h>
#include <omp.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include <zlib.h>

#include <algorithm>
#include <atomic>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "mp2_common.h"

struct mp2_bam {
    std::vector<std::string> names;
    std::vector<int64_t> lens;
    std::vector<int64_t> ev_off;
    uint32_t *starts = nullptr, *ends = nullptr;
    size_t cap_s = 0, cap_e = 0;  // bytes of the two blocks (they go back to the pinned pool)
    bool pinned = false, pin_s = false, pin_e = false;
    int64_t n_events = 0, n_records = 0, n_kept = 0;
    uint32_t max_span = 0, max_block = 0, max_disorder = 0;
    std::string header_text;
};

namespace mp2 {

struct Block {
    int64_t coff;   // offset of the block in the file
    uint32_t csize; // whole block size
    uint32_t usize; // ISIZE
};

static inline uint32_t rd32(const uint8_t *p) { uint32_t v; memcpy(&v, p, 4); return v; }
static inline int32_t rds32(const uint8_t *p) { int32_t v; memcpy(&v, p, 4); return v; }
static inline uint16_t rd16(const uint8_t *p) { uint16_t v; memcpy(&v, p, 2); return v; }

// Index of the BGZF blocks (SAM spec 4.1: gzip member with a 'BC' extra subfield holding BSIZE-1).
static int index_blocks(const uint8_t *f, int64_t size, std::vector<Block> &blocks, const char *path) {
    int64_t o = 0;
    while (o < size) {
        if (size - o < 28) return fail(MP2_EIO, "%s: truncated BGZF block header at %lld", path, (long long)o);
        const uint8_t *h = f + o;
        if (h[0] != 31 || h[1] != 139 || h[2] != 8 || !(h[3] & 4))
            return fail(MP2_EIO, "%s: not a BGZF file (bad gzip member at %lld)", path, (long long)o);
        const uint32_t xlen = rd16(h + 10);
        int64_t bsize = -1;
        for (uint32_t x = 0; x + 4 <= xlen;) {
            const uint8_t *s = h + 12 + x;
            const uint32_t slen = rd16(s + 2);
            if (s[0] == 'B' && s[1] == 'C' && slen == 2) bsize = (int64_t)rd16(s + 4) + 1;
            x += 4 + slen;
        }
        if (bsize < 0 || o + bsize > size)
            return fail(MP2_EIO, "%s: BGZF block at %lld has no BC field or runs past the end", path, (long long)o);
        Block b;
        b.coff = o;
        b.csize = (uint32_t)bsize;
        b.usize = rd32(f + o + bsize - 4);
        if (b.usize > 65536) return fail(MP2_EIO, "%s: BGZF block larger than 64 KiB", path);
        blocks.push_back(b);
        o += bsize;
    }
    return MP2_OK;
}

static bool inflate_block(const uint8_t *f, const Block &b, uint8_t *dst) {
    if (b.usize == 0) return true;
    const uint32_t xlen = rd16(f + b.coff + 10);
    const uint8_t *src = f + b.coff + 12 + xlen;
    const int64_t clen = (int64_t)b.csize - 12 - xlen - 8;
    if (clen < 0) return false;
    z_stream zs;
    memset(&zs, 0, sizeof zs);
    if (inflateInit2(&zs, -15) != Z_OK) return false;
    zs.next_in = const_cast<Bytef *>(src);
    zs.avail_in = (uInt)clen;
    zs.next_out = dst;
    zs.avail_out = b.usize;
    const int rc = inflate(&zs, Z_FINISH);
    const bool ok = (rc == Z_STREAM_END) && zs.total_out == b.usize;
    inflateEnd(&zs);
    return ok;
}

// Aux field `t0 t1` of a record: pointer to its value and its type, or NULL.  *bad is set when the aux
// area is malformed.
static const uint8_t *find_aux(const uint8_t *a, const uint8_t *end, char t0, char t1, uint8_t &type, bool &bad) {
    while (a + 3 <= end) {
        const uint8_t k0 = a[0], k1 = a[1], ty = a[2];
        a += 3;
        size_t sz = 0;
        switch (ty) {
        case 'A': case 'c': case 'C': sz = 1; break;
        case 's': case 'S': sz = 2; break;
        case 'i': case 'I': case 'f': sz = 4; break;
        case 'Z': case 'H': {
            const uint8_t *z = (const uint8_t *)memchr(a, 0, end - a);
            if (!z) { bad = true; return nullptr; }
            sz = (z - a) + 1;
            break;
        }
        case 'B': {
            if (a + 5 > end) { bad = true; return nullptr; }
            const uint8_t st = a[0];
            const uint32_t cnt = rd32(a + 1);
            const size_t es = (st == 'c' || st == 'C') ? 1 : (st == 's' || st == 'S') ? 2 : 4;
            sz = 5 + es * (size_t)cnt;
            break;
        }
        default: bad = true; return nullptr;
        }
        if (a + sz > end) { bad = true; return nullptr; }
        if (k0 == (uint8_t)t0 && k1 == (uint8_t)t1) { type = ty; return a; }
        a += sz;
    }
    return nullptr;
}

// One record -> alignment blocks, straight from the BAM wire format.
// Returns the number of M/=/X blocks (0 = dropped), -1 = NM tag missing, -2 = malformed.  The first block
// goes to (s0, e0), the others to more_s / more_e (when given); `pos` is the record's position.
// A read with more than 65535 CIGAR operations keeps its real CIGAR in the CG:B,I tag behind a
// placeholder "<l_seq>S<ref_len>N" (SAM spec 4.2.2); htslib, which the reference reads BAMs with, swaps it in
// when the record is read, and so does this function.
static int record_blocks(const uint8_t *r, uint32_t block_size, float min_identity, int32_t &tid, int32_t &pos,
                         uint32_t &s0, uint32_t &e0, uint32_t *more_s, uint32_t *more_e, int max_more) {
    if (block_size < 32) return -2;
    tid = rds32(r);
    pos = rds32(r + 4);
    const uint32_t l_read_name = r[8];
    uint32_t n_cigar = rd16(r + 12);
    const uint32_t flag = rd16(r + 14);
    const uint32_t l_seq = rd32(r + 16);
    const uint64_t fixed = 32ull + l_read_name + 4ull * n_cigar + (l_seq + 1) / 2 + l_seq;
    if (fixed > block_size) return -2;
    if (flag & 0x100) return 0;  // secondary      (src/coverage.rs:104)
    if (flag & 0x800) return 0;  // supplementary  (src/coverage.rs:105)
    if ((flag & 0x4) || tid < 0 || pos < 0) return 0;  // unmapped: contributes nothing
    const uint8_t *cig = r + 32 + l_read_name;
    const uint8_t *aux = r + fixed, *end = r + block_size;
    if (n_cigar == 2) {
        const uint32_t c0 = rd32(cig), c1 = rd32(cig + 4);
        if ((c0 & 0xF) == 4 && (c0 >> 4) == l_seq && (c1 & 0xF) == 3) {  // placeholder: real CIGAR in CG:B,I
            uint8_t ty = 0;
            bool bad = false;
            const uint8_t *v = find_aux(aux, end, 'C', 'G', ty, bad);
            if (bad) return -2;
            if (v && ty == 'B' && (v[0] == 'I' || v[0] == 'i')) {
                n_cigar = rd32(v + 1);
                cig = v + 5;
            }
        }
    }
    if (min_identity > 0.0f) {
        // NM aux tag (required, as in coverm: the reference panics without it)
        uint8_t ty = 0;
        bool bad = false;
        const uint8_t *v = find_aux(aux, end, 'N', 'M', ty, bad);
        if (bad) return -2;
        if (!v) return -1;
        int64_t nm;
        switch (ty) {
        case 'c': nm = (int8_t)v[0]; break;
        case 'C': nm = v[0]; break;
        case 's': nm = (int16_t)rd16(v); break;
        case 'S': nm = rd16(v); break;
        case 'i': nm = rds32(v); break;
        case 'I': nm = rd32(v); break;
        default: return -2;
        }
        if (nm < 0) return -1;
        uint32_t aligned = 0;
        for (uint32_t i = 0; i < n_cigar; i++) {
            const uint32_t c = rd32(cig + 4 * i), op = c & 0xF, l = c >> 4;
            if (op == 0 || op == 1 || op == 7 || op == 8) aligned += l;  // M I = X
        }
        volatile float q = (float)nm / (float)aligned;
        volatile float ident = 1.0f - q;
        if (!(ident >= min_identity)) return 0;
    }
    int nb = 0;
    uint32_t cursor = (uint32_t)pos;
    for (uint32_t i = 0; i < n_cigar; i++) {
        const uint32_t c = rd32(cig + 4 * i), op = c & 0xF, l = c >> 4;
        if (op == 0 || op == 7 || op == 8) {  // M = X
            if (nb == 0) { s0 = cursor; e0 = cursor + l; }
            else if (nb - 1 < max_more) { more_s[nb - 1] = cursor; more_e[nb - 1] = cursor + l; }
            nb++;
            cursor += l;
        } else if (op == 2 || op == 3) {  // D N
            cursor += l;
        }  // I S H P: nothing
    }
    return nb;
}

// ---- pinned host memory, pooled ------------------------------------------------------------------------
// Page-locking is slow (tens of ms for a sample's events), so the blocks of a closed sample go back to a
// small process-wide pool and the next sample decodes straight into them: in a run over many BAMs the
// events are written once, into pinned memory, and copied to the device from there.
struct PinBlock { void *p; size_t bytes; bool pinned; };
static std::mutex g_pin_mu;
static std::vector<PinBlock> g_pin_pool;
constexpr size_t PIN_POOL_MAX = 8;

static PinBlock pin_get(size_t bytes) {
    if (bytes < 4096) bytes = 4096;
    {
        std::lock_guard<std::mutex> lk(g_pin_mu);
        int best = -1;
        for (size_t i = 0; i < g_pin_pool.size(); i++)
            if (g_pin_pool[i].bytes >= bytes && (best < 0 || g_pin_pool[i].bytes < g_pin_pool[best].bytes)) best = (int)i;
        if (best >= 0) {
            PinBlock b = g_pin_pool[best];
            g_pin_pool.erase(g_pin_pool.begin() + best);
            return b;
        }
    }
    PinBlock b{nullptr, bytes, false};
    if (cudaHostAlloc(&b.p, bytes, cudaHostAllocPortable) == cudaSuccess) {
        b.pinned = true;
    } else {
        (void)cudaGetLastError();
        b.p = malloc(bytes);
    }
    return b;
}

static void pin_put(PinBlock b) {
    if (!b.p) return;
    {
        std::lock_guard<std::mutex> lk(g_pin_mu);
        if (b.pinned && g_pin_pool.size() < PIN_POOL_MAX) {
            g_pin_pool.push_back(b);
            return;
        }
    }
    if (b.pinned) cudaFreeHost(b.p);
    else free(b.p);
}

// growable uint32 array in pooled pinned memory
struct PinVec {
    PinBlock blk{nullptr, 0, false};
    size_t n = 0;
    uint32_t *data() { return (uint32_t *)blk.p; }
    // `expect` = how many elements the array will probably hold in the end (0 = unknown): page-locking is the
    // expensive part, so the first block is taken at that size and growth is the exception
    bool reserve(size_t want, int threads, size_t expect = 0) {
        if (want * 4 <= blk.bytes) return true;
        size_t cap = std::max(std::max(want, expect) * 4, blk.bytes + blk.bytes / 2);
        PinBlock nb = pin_get(cap);
        if (!nb.p) return false;
        if (n) {
            const size_t bytes = n * 4, grain = 1u << 20;
            const int64_t parts = (int64_t)((bytes + grain - 1) / grain);
#pragma omp parallel for schedule(static) num_threads(threads)
            for (int64_t i = 0; i < parts; i++)
                memcpy((char *)nb.p + i * grain, (const char *)blk.p + i * grain, std::min(grain, bytes - (size_t)i * grain));
        }
        pin_put(blk);
        blk = nb;
        return true;
    }
    ~PinVec() { pin_put(blk); }
};

}  // namespace mp2

using namespace mp2;

extern "C" int mp2_bam_open(const char *path, float min_identity, int threads, mp2_bam **out) {
    if (!path || !out) return fail(MP2_EINVAL, "NULL argument");
    *out = nullptr;
    if (threads < 1) threads = 1;
    const int fd = open(path, O_RDONLY);
    if (fd < 0) return fail(MP2_EIO, "cannot open %s", path);
    struct stat sb;
    if (fstat(fd, &sb) != 0 || sb.st_size <= 0) {
        close(fd);
        return fail(MP2_EIO, "cannot stat %s (or it is empty)", path);
    }
    const int64_t size = sb.st_size;
    const uint8_t *f = (const uint8_t *)mmap(nullptr, size, PROT_READ, MAP_PRIVATE, fd, 0);
    close(fd);
    if (f == MAP_FAILED) return fail(MP2_EIO, "cannot mmap %s", path);
    struct Unmap {
        const uint8_t *p; int64_t n;
        ~Unmap() { munmap((void *)p, n); }
    } unmap{f, size};
    madvise((void *)f, size, MADV_SEQUENTIAL);

    const bool trace = getenv("MP2_TRACE") != nullptr;
    double t_idx = omp_get_wtime(), t_inf = 0, t_hop = 0, t_flt = 0, t_mrg = 0, t_res = 0, t_x;
    std::vector<Block> blocks;
    MP2_TRY(index_blocks(f, size, blocks, path));
    t_idx = omp_get_wtime() - t_idx;

    mp2_bam *b = new mp2_bam();
    struct Guard {
        mp2_bam *b;
        ~Guard() { if (b) mp2_bam_close(b); }
    } guard{b};

    // uncompressed window (leftover + batch); grown with realloc so that nothing is ever zero-filled
    struct RawBuf {
        uint8_t *p = nullptr;
        size_t cap = 0;
        ~RawBuf() { free(p); }
        bool reserve(size_t n) {
            if (n <= cap) return true;
            const size_t c = n + n / 4 + 4096;
            uint8_t *q = (uint8_t *)realloc(p, c);
            if (!q) return false;
            p = q; cap = c;
            return true;
        }
        uint8_t *data() { return p; }
    } buf;
    PinVec ev_s, ev_e;  // accepted blocks in BAM order, decoded straight into (pooled) pinned memory
    std::vector<int64_t> per_tid;
    int64_t BATCH = 64ll << 20;  // uncompressed bytes inflated per round (MP2_BAM_BATCH_KB: test switch)
    if (const char *e = getenv("MP2_BAM_BATCH_KB")) {
        const long kb = atol(e);
        if (kb > 0) BATCH = (int64_t)kb << 10;
    }
    int64_t total_usize = 0, done_usize = 0;  // uncompressed bytes of the file / inflated so far
    for (const Block &blk : blocks) total_usize += blk.usize;
    size_t bi = 0;
    int64_t have = 0;   // valid bytes in buf
    bool header_done = false;
    int64_t last_tid = -1, last_pos = -1;
    bool unsorted_pos = false;
    std::vector<int64_t> rec_off;
    std::vector<int32_t> rec_tid;
    std::vector<int32_t> rec_n;
    std::vector<uint32_t> rec_s, rec_e;
    std::vector<int32_t> rec_pos;

    while (bi < blocks.size() || have > 0) {
        // ---- inflate the next batch of blocks behind the leftover bytes ----
        size_t bj = bi;
        int64_t add = 0;
        while (bj < blocks.size() && (add == 0 || add + blocks[bj].usize <= BATCH)) add += blocks[bj++].usize;
        t_x = omp_get_wtime();
        if (!buf.reserve((size_t)(have + add))) return fail(MP2_ENOMEM, "%s: out of host memory for the inflate window", path);
        t_res += omp_get_wtime() - t_x;
        t_x = omp_get_wtime();
        std::vector<int64_t> uoff(bj - bi + 1, have);
        for (size_t k = bi; k < bj; k++) uoff[k - bi + 1] = uoff[k - bi] + blocks[k].usize;
        std::atomic<int> bad{0};
#pragma omp parallel for schedule(dynamic, 16) num_threads(threads)
        for (int64_t k = (int64_t)bi; k < (int64_t)bj; k++)
            if (!inflate_block(f, blocks[k], buf.data() + uoff[k - bi])) bad = 1;
        if (bad) return fail(MP2_EIO, "%s: corrupt BGZF block (inflate failed)", path);
        t_inf += omp_get_wtime() - t_x;
        const bool last_batch = bj == blocks.size();
        bi = bj;
        have += add;
        done_usize += add;
        const uint8_t *p = buf.data();
        int64_t o = 0;
        // ---- header ----
        if (!header_done) {
            if (have < 12) {
                if (last_batch) return fail(MP2_EIO, "%s: truncated BAM header", path);
                continue;
            }
            if (memcmp(p, "BAM\1", 4) != 0) return fail(MP2_EIO, "%s: not a BAM file (bad magic)", path);
            const int64_t l_text = rds32(p + 4);
            if (l_text < 0 || 12 + l_text > have) {
                if (last_batch) return fail(MP2_EIO, "%s: truncated BAM header", path);
                continue;  // header larger than one batch: keep accumulating
            }
            const int64_t n_ref = rds32(p + 8 + l_text);
            int64_t q = 12 + l_text;
            bool complete = n_ref >= 0;
            std::vector<std::string> names;
            std::vector<int64_t> lens;
            for (int64_t r = 0; complete && r < n_ref; r++) {
                if (q + 4 > have) { complete = false; break; }
                const int64_t l_name = rds32(p + q);
                if (l_name < 1 || q + 4 + l_name + 4 > have) { complete = false; break; }
                names.emplace_back((const char *)p + q + 4, (size_t)l_name - 1);
                lens.push_back((int64_t)rd32(p + q + 4 + l_name));
                q += 8 + l_name;
            }
            if (!complete) {
                if (last_batch) return fail(MP2_EIO, "%s: truncated BAM reference list", path);
                continue;
            }
            b->header_text.assign((const char *)p + 8, (size_t)l_text);
            b->names.swap(names);
            b->lens.swap(lens);
            per_tid.assign(b->names.size(), 0);
            header_done = true;
            o = q;
        }
        // ---- record boundaries (sequential hop), then parallel filter ----
        t_x = omp_get_wtime();
        // (the only serial pass over the bytes: it reads 4 bytes per record, a few records ahead are
        // prefetched because the stride is the record size and the hardware prefetcher does not follow it)
        rec_off.clear();
        {
            int64_t stride = 256;
            while (o + 4 <= have) {
                const int64_t bs = rd32(p + o);
                if (o + 4 + bs > have) break;
                rec_off.push_back(o);
                stride = 4 + bs;
                if (o + 12 * stride + 4 <= have) {
                    __builtin_prefetch(p + o + 8 * stride);
                    __builtin_prefetch(p + o + 12 * stride);
                }
                o += stride;
            }
        }
        t_hop += omp_get_wtime() - t_x;
        t_x = omp_get_wtime();
        const int64_t nr = (int64_t)rec_off.size();
        rec_tid.resize(nr);
        rec_n.resize(nr);
        rec_s.resize(nr);
        rec_e.resize(nr);
        rec_pos.resize(nr);
        std::atomic<int> err{0};
#pragma omp parallel for schedule(static) num_threads(threads)
        for (int64_t i = 0; i < nr; i++) {
            const uint8_t *r = p + rec_off[i];
            int32_t tid = -1, rpos = 0;
            uint32_t s0 = 0, e0 = 0;
            const int k = record_blocks(r + 4, rd32(r), min_identity, tid, rpos, s0, e0, nullptr, nullptr, 0);
            if (k < 0) err = k == -1 ? 1 : 2;
            rec_tid[i] = tid;
            rec_pos[i] = rpos;
            rec_n[i] = k;
            rec_s[i] = s0;
            rec_e[i] = e0;
        }
        if (err == 1)
            return fail(MP2_EFORMAT, "%s: a mapped record has no NM tag (needed for the identity filter; "
                        "the reference panics here)", path);
        if (err) return fail(MP2_EIO, "%s: malformed BAM record", path);
        t_flt += omp_get_wtime() - t_x;
        t_x = omp_get_wtime();
        // ---- merge: serial scan of the per-record block counts (order, totals), parallel fill ----
        std::vector<int64_t> ev_pos((size_t)nr + 1);
        {
            int64_t pos = (int64_t)ev_s.n;
            for (int64_t i = 0; i < nr; i++) {
                ev_pos[i] = pos;
                const int k = rec_n[i];
                if (k <= 0) continue;
                const int64_t tid = rec_tid[i];
                if (tid >= (int64_t)b->names.size())
                    return fail(MP2_EIO, "%s: record refers to an unknown reference", path);
                if (tid < last_tid)
                    return fail(MP2_EFORMAT, "%s: records are not sorted by reference (SO:coordinate required)", path);
                if (tid != last_tid) last_pos = -1;
                if (rec_pos[i] < last_pos) unsorted_pos = true;
                last_pos = rec_pos[i];
                last_tid = tid;
                b->n_kept++;
                per_tid[tid] += k;
                pos += k;
            }
            ev_pos[nr] = pos;
            b->n_records += nr;
            // blocks per uncompressed byte so far, extrapolated to the whole file (+ 6 %)
            const double seen = (double)(done_usize - (have - o));
            const size_t expect = seen > 0 ? (size_t)((double)pos * ((double)total_usize / seen) * 1.06) + 1024 : 0;
            if (!ev_s.reserve((size_t)pos, threads, expect) || !ev_e.reserve((size_t)pos, threads, expect))
                return fail(MP2_ENOMEM, "%s: out of host memory for %lld alignment blocks", path, (long long)pos);
            ev_s.n = ev_e.n = (size_t)pos;
        }
        uint32_t span_max = b->max_span, block_max = b->max_block, dis_max = b->max_disorder;
        uint32_t *const out_s = ev_s.data(), *const out_e = ev_e.data();
#pragma omp parallel for schedule(static) num_threads(threads) reduction(max : span_max, block_max, dis_max)
        for (int64_t i = 0; i < nr; i++) {
            const int k = rec_n[i];
            if (k <= 0) continue;
            const int64_t at = ev_pos[i];
            const uint32_t rpos = (uint32_t)rec_pos[i];
            if (k == 1) {
                out_s[at] = rec_s[i];
                out_e[at] = rec_e[i];
            } else {
                int32_t t2 = 0, p2 = 0; uint32_t s0 = 0, e0 = 0;
                const uint8_t *r = p + rec_off[i];
                record_blocks(r + 4, rd32(r), min_identity, t2, p2, s0, e0, out_s + at + 1, out_e + at + 1, k - 1);
                out_s[at] = s0; out_e[at] = e0;
            }
            for (int j = 0; j < k; j++) {
                const uint32_t blk = out_e[at + j] - out_s[at + j], dis = out_s[at + j] - rpos;
                block_max = blk > block_max ? blk : block_max;
                dis_max = dis > dis_max ? dis : dis_max;   // how far this start can lie below a later read's
            }
            const uint32_t span = out_e[at + k - 1] - rpos;  // the whole read footprint
            span_max = span > span_max ? span : span_max;
        }
        b->max_block = block_max;
        b->max_disorder = dis_max;
        b->max_span = span_max;
        t_mrg += omp_get_wtime() - t_x;
        // ---- keep the partial record for the next batch ----
        const int64_t left = have - o;
        if (left > 0 && last_batch) {
            return fail(MP2_EIO, "%s: truncated BAM record at end of file", path);
        }
        if (left > 0) memmove(buf.data(), buf.data() + o, left);
        have = left;
        if (last_batch) break;
    }
    if (!header_done) return fail(MP2_EIO, "%s: no BAM header", path);

    const int64_t n = (int64_t)b->names.size();
    b->ev_off.assign(n + 1, 0);
    for (int64_t i = 0; i < n; i++) b->ev_off[i + 1] = b->ev_off[i] + per_tid[i];
    b->n_events = (int64_t)ev_s.n;
    if (!ev_s.reserve(4, threads) || !ev_e.reserve(4, threads)) return fail(MP2_ENOMEM, "out of host memory");
    // the sample keeps the two blocks (mp2_bam_close gives them back to the pool)
    b->starts = ev_s.data(); b->cap_s = ev_s.blk.bytes;
    b->ends = ev_e.data(); b->cap_e = ev_e.blk.bytes;
    b->pinned = ev_s.blk.pinned && ev_e.blk.pinned;
    b->pin_s = ev_s.blk.pinned; b->pin_e = ev_e.blk.pinned;
    ev_s.blk = PinBlock{nullptr, 0, false};
    ev_e.blk = PinBlock{nullptr, 0, false};
    if (unsorted_pos) b->max_disorder = 0xFFFFFFFFu;  // positions go back inside a reference: general path
    if (trace)
        fprintf(stderr, "[mp2_bam] index %.3f resize %.3f inflate %.3f hop %.3f filter %.3f merge %.3f s (%d threads)\n",
                t_idx, t_res, t_inf, t_hop, t_flt, t_mrg, threads);
    guard.b = nullptr;
    *out = b;
    return MP2_OK;
}

extern "C" {

int64_t mp2_bam_n_contigs(const mp2_bam *b) { return b ? (int64_t)b->names.size() : 0; }
int64_t mp2_bam_n_events(const mp2_bam *b) { return b ? b->n_events : 0; }
int64_t mp2_bam_n_records(const mp2_bam *b) { return b ? b->n_records : 0; }
int64_t mp2_bam_n_kept(const mp2_bam *b) { return b ? b->n_kept : 0; }
uint32_t mp2_bam_max_span(const mp2_bam *b) { return b ? b->max_span : 0; }
uint32_t mp2_bam_max_block(const mp2_bam *b) { return b ? b->max_block : 0; }
uint32_t mp2_bam_max_disorder(const mp2_bam *b) { return b ? b->max_disorder : 0; }
const char *mp2_bam_contig_name(const mp2_bam *b, int64_t i) {
    return (b && i >= 0 && i < (int64_t)b->names.size()) ? b->names[i].c_str() : nullptr;
}
const int64_t *mp2_bam_contig_len(const mp2_bam *b) { return b ? b->lens.data() : nullptr; }
const int64_t *mp2_bam_ev_off(const mp2_bam *b) { return b ? b->ev_off.data() : nullptr; }
const uint32_t *mp2_bam_starts(const mp2_bam *b) { return b ? b->starts : nullptr; }
const uint32_t *mp2_bam_ends(const mp2_bam *b) { return b ? b->ends : nullptr; }
const char *mp2_bam_header_text(const mp2_bam *b, int64_t *len) {
    if (!b) return nullptr;
    if (len) *len = (int64_t)b->header_text.size();
    return b->header_text.data();
}
void mp2_bam_close(mp2_bam *b) {
    if (!b) return;
    mp2::pin_put(mp2::PinBlock{b->starts, b->cap_s, b->pin_s});
    mp2::pin_put(mp2::PinBlock{b->ends, b->cap_e, b->pin_e});
    delete b;
}

}  // extern "C"
