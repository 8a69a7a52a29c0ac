// Host-side FASTA ingest: plain or gzip (zlib), records concatenated into ONE byte stream plus
// offsets -- the layout mp2_tnf takes.  Replaces tools.read_fasta under core.Mag
// (/root/reference/magpurify2/tools.py:210-240, magpurify2/core.py:35-50): record id = first
// word of the header with '~' -> '_', description = the whole header, sequence = the lines joined as
// Biopython's SimpleFastaParser joins them (trailing white space stripped, ' ' and CR removed).
#include <sys/stat.h>
#include <zlib.h>

#include <cstring>
#include <string>
#include <vector>

#include "mp2_common.h"

struct mp2_fasta {
    std::vector<std::string> ids, descs;
    std::vector<int64_t> offsets;
    uint8_t *bases = nullptr;
    bool pinned = false;
    int64_t n_bases = 0;
};

namespace mp2 {

static int parse_fasta(const uint8_t *p, int64_t size, mp2_fasta *f) {
    // The payload is never larger than the file: one allocation, one pass.  Plain host memory: the
    // drivers concatenate the MAGs of a run into one stream, which is what gets staged to the device.
    const size_t bytes = (size_t)size + 64;
    void *mem = malloc(bytes);
    if (!mem) return fail(MP2_ENOMEM, "out of host memory for %lld bytes of FASTA", (long long)size);
    f->bases = (uint8_t *)mem;
    f->pinned = false;
    int64_t w = 0, i = 0;
    bool seen = false;
    while (i < size) {
        const uint8_t *nl = (const uint8_t *)memchr(p + i, '\n', size - i);
        const int64_t e = nl ? nl - p : size;  // line = [i, e)
        if (p[i] == '>') {
            int64_t le = e;
            while (le > i + 1 && (p[le - 1] == '\r' || p[le - 1] == ' ' || p[le - 1] == '\t')) le--;
            std::string desc((const char *)p + i + 1, (size_t)(le - i - 1));
            size_t k = 0;
            while (k < desc.size() && !(desc[k] == ' ' || desc[k] == '\t')) k++;
            std::string id = desc.substr(0, k);
            for (auto &ch : id)
                if (ch == '~') ch = '_';  // tools.py:239
            if (seen) f->offsets.push_back(w);
            else f->offsets.push_back(0);
            f->ids.push_back(id);
            f->descs.push_back(desc);
            seen = true;
        } else if (seen) {
            // sequence line, as Biopython's SimpleFastaParser builds it (Bio/SeqIO/FastaIO.py, under
            // tools.read_fasta): trailing white space stripped, then every ' ' and '\r' removed; any other
            // byte is kept.  Lines are almost always clean, so look for the two bytes with a branch-free
            // (vectorisable) scan and copy the line in one go.
            int64_t le = e;
            while (le > i && (p[le - 1] == ' ' || (p[le - 1] >= 0x09 && p[le - 1] <= 0x0D))) le--;
            unsigned dirty = 0;
            for (int64_t k = i; k < le; k++) dirty |= (unsigned)(p[k] == ' ') | (unsigned)(p[k] == '\r');
            if (!dirty) {
                memcpy(f->bases + w, p + i, (size_t)(le - i));
                w += le - i;
            } else {
                for (int64_t k = i; k < le; k++)
                    if (p[k] != ' ' && p[k] != '\r') f->bases[w++] = p[k];
            }
        }
        i = e + 1;
    }
    f->offsets.push_back(w);
    if (!seen) f->offsets.assign(1, 0);
    f->n_bases = w;
    memset(f->bases + w, 0, 64);
    return MP2_OK;
}

}  // namespace mp2

using namespace mp2;

extern "C" void mp2_fasta_close(mp2_fasta *f) {
    if (!f) return;
    if (f->bases) {
        if (f->pinned) cudaFreeHost(f->bases);
        else free(f->bases);
    }
    delete f;
}

extern "C" int mp2_fasta_open_mem(const uint8_t *data, int64_t size, mp2_fasta **out) {
    if (!out || (size > 0 && !data) || size < 0) return fail(MP2_EINVAL, "bad argument");
    *out = nullptr;
    mp2_fasta *f = new mp2_fasta();
    const int rc = parse_fasta(data, size, f);
    if (rc != MP2_OK) {
        mp2_fasta_close(f);
        return rc;
    }
    *out = f;
    return MP2_OK;
}

extern "C" int mp2_fasta_open(const char *path, mp2_fasta **out) {
    if (!path || !out) return fail(MP2_EINVAL, "NULL argument");
    *out = nullptr;
    gzFile g = gzopen(path, "rb");  // transparent for files that are not gzip
    if (!g) return fail(MP2_EIO, "cannot open %s", path);
    gzbuffer(g, 1 << 20);
    // realloc'ed window (never zero-filled); a plain file is read in one go, a gzip one grows by doubling
    struct Raw {
        uint8_t *p = nullptr;
        ~Raw() { free(p); }
    } buf;
    size_t cap = 0, have = 0;
    {
        struct stat sb;
        if (stat(path, &sb) == 0 && sb.st_size > 0) cap = (size_t)sb.st_size + (1u << 16);
    }
    if (cap < (1u << 20)) cap = 1u << 20;
    buf.p = (uint8_t *)malloc(cap);
    if (!buf.p) { gzclose(g); return fail(MP2_ENOMEM, "%s: out of host memory", path); }
    for (;;) {
        if (cap - have < (1u << 16)) {
            uint8_t *q = (uint8_t *)realloc(buf.p, cap * 2);
            if (!q) { gzclose(g); return fail(MP2_ENOMEM, "%s: out of host memory", path); }
            buf.p = q;
            cap *= 2;
        }
        const int got = gzread(g, buf.p + have, (unsigned)std::min<size_t>(cap - have, 1u << 30));
        if (got < 0) {
            int e = 0;
            const char *msg = gzerror(g, &e);
            std::string m = msg ? msg : "read error";
            gzclose(g);
            return fail(MP2_EIO, "%s: %s", path, m.c_str());
        }
        if (got == 0) break;
        have += (size_t)got;
    }
    gzclose(g);
    return mp2_fasta_open_mem(buf.p, (int64_t)have, out);
}

extern "C" {
int64_t mp2_fasta_n(const mp2_fasta *f) { return f ? (int64_t)f->ids.size() : 0; }
const uint8_t *mp2_fasta_bases(const mp2_fasta *f) { return f ? f->bases : nullptr; }
const int64_t *mp2_fasta_offsets(const mp2_fasta *f) { return f ? f->offsets.data() : nullptr; }
const char *mp2_fasta_id(const mp2_fasta *f, int64_t i) {
    return (f && i >= 0 && i < (int64_t)f->ids.size()) ? f->ids[i].c_str() : nullptr;
}
const char *mp2_fasta_description(const mp2_fasta *f, int64_t i) {
    return (f && i >= 0 && i < (int64_t)f->descs.size()) ? f->descs[i].c_str() : nullptr;
}
}
