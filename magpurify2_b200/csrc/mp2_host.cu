// Host-buffer entry points: stage H2D/D2H around the device kernels.
//
// Pageable user memory is copied through a small ring of pinned buffers (parallel memcpy into the
// ring, cudaMemcpyAsync out of it) so the PCIe copy of slice k+1 overlaps the kernels of slice k;
// memory that is already pinned/registered is copied directly.
#include <omp.h>

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <vector>

#include "mp2_common.h"

namespace mp2 {

int tnf_launch_range(const uint8_t *d_bases, const int64_t *d_offsets, int64_t n, int64_t n_bases,
                     int64_t chunk_begin, int64_t chunk_end, uint32_t *d_counts,
                     unsigned long long *d_gc_counts, cudaStream_t stream, bool zero_rows);
int64_t tnf_num_chunks(const void *d_bases, int64_t n_bases);
int64_t tnf_chunks_ready(const void *d_bases, int64_t done, int64_t n_bases);
int tnf_valid_from_runs(const int64_t *d_runs, int64_t n_runs, int64_t n_bases, uint32_t *d_valid, cudaStream_t stream);
int64_t tnf_packed_num_chunks(int64_t n_bases);
int64_t tnf_packed_bases_done(int64_t chunks, int64_t n_bases);
int64_t tnf_packed_chunks_ready(int64_t done, int64_t n_bases);
int tnf_packed_launch_range(const uint32_t *d_codes, const uint32_t *d_valid, const int64_t *d_offsets, int64_t n,
                            int64_t n_bases, int64_t chunk_begin, int64_t chunk_end, uint32_t *d_counts,
                            unsigned long long *d_gc_counts, cudaStream_t stream);
int tnf_finalize(const uint32_t *d_counts, const unsigned long long *d_gc_counts,
                 const int64_t *d_offsets, int64_t n, float *d_rel, float *d_gc,
                 cudaStream_t stream);

static size_t slice_bytes() {
    static size_t v = 0;
    if (!v) {
        const char *e = getenv("MP2_SLICE_MB");
        long mb = e ? atol(e) : 0;
        v = (size_t)(mb > 0 ? mb : 32) << 20;
    }
    return v;
}
#define SLICE (slice_bytes())
constexpr size_t RING_SLICE = 16u << 20;
constexpr int NRING = 4;

// One ring per device: the slots are plain (portable) pinned memory, but the events that guard them belong to
// the device they were created on, and threads driving different GPUs must not wait for each other.
struct Ring {
    uint8_t *buf[NRING] = {nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t freed[NRING];
    bool ready = false;
    std::mutex mu;
};
constexpr int MAX_DEV = 64;
static Ring g_rings[MAX_DEV];

static int ring_of_current_device(Ring **out) {
    int dev = 0;
    MP2_CUDA(cudaGetDevice(&dev));
    if (dev < 0 || dev >= MAX_DEV) return fail(MP2_EINVAL, "device index %d out of range", dev);
    *out = &g_rings[dev];
    return MP2_OK;
}

static int ring_init(Ring &r) {  // called with r.mu held
    if (r.ready) return MP2_OK;
    for (int i = 0; i < NRING; i++) {
        MP2_CUDA(cudaHostAlloc((void **)&r.buf[i], RING_SLICE, cudaHostAllocPortable));
        MP2_CUDA(cudaEventCreateWithFlags(&r.freed[i], cudaEventDisableTiming));
    }
    r.ready = true;
    return MP2_OK;
}

static bool is_pinned(const void *p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        (void)cudaGetLastError();
        return false;
    }
    return a.type == cudaMemoryTypeHost;
}

static void par_memcpy(void *dst, const void *src, size_t bytes) {
    const size_t grain = 1u << 20;
    if (bytes <= 2 * grain) {
        memcpy(dst, src, bytes);
        return;
    }
    const int64_t parts = (int64_t)((bytes + grain - 1) / grain);
    int nt = std::min<int64_t>(std::min(parts, (int64_t)omp_get_max_threads()), 16);
#pragma omp parallel for schedule(static) num_threads(nt)
    for (int64_t i = 0; i < parts; i++) {
        size_t o = (size_t)i * grain;
        memcpy((char *)dst + o, (const char *)src + o, std::min(grain, bytes - o));
    }
}

// H2D of a whole host array (pinned: one async copy; pageable: through the ring).
// `after_slice(bytes_done)` is called after each slice has been enqueued on `copy`.
template <class F>
static int upload(void *d_dst, const void *h_src, size_t bytes, cudaStream_t copy, F after_slice) {
    if (bytes == 0) return after_slice((size_t)0);
    if (is_pinned(h_src)) {
        for (size_t o = 0; o < bytes; o += SLICE) {
            size_t len = std::min(SLICE, bytes - o);
            MP2_CUDA(cudaMemcpyAsync((char *)d_dst + o, (const char *)h_src + o, len,
                                     cudaMemcpyHostToDevice, copy));
            MP2_TRY(after_slice(o + len));
        }
        return MP2_OK;
    }
    Ring *ring = nullptr;
    MP2_TRY(ring_of_current_device(&ring));
    Ring &g_ring = *ring;
    std::lock_guard<std::mutex> lk(g_ring.mu);
    MP2_TRY(ring_init(g_ring));
    int slot = 0;
    for (size_t o = 0; o < bytes; o += RING_SLICE, slot = (slot + 1) % NRING) {
        size_t len = std::min(RING_SLICE, bytes - o);
        MP2_CUDA(cudaEventSynchronize(g_ring.freed[slot]));  // previous user of the slot is done
        par_memcpy(g_ring.buf[slot], (const char *)h_src + o, len);
        MP2_CUDA(cudaMemcpyAsync((char *)d_dst + o, g_ring.buf[slot], len, cudaMemcpyHostToDevice,
                                 copy));
        MP2_CUDA(cudaEventRecord(g_ring.freed[slot], copy));
        MP2_TRY(after_slice(o + len));
    }
    return MP2_OK;
}

static int upload(void *d_dst, const void *h_src, size_t bytes, cudaStream_t copy) {
    return upload(d_dst, h_src, bytes, copy, [](size_t) { return MP2_OK; });
}

// D2H into user memory; caller synchronises the stream afterwards.
static int download(void *h_dst, const void *d_src, size_t bytes, cudaStream_t s) {
    if (bytes == 0) return MP2_OK;
    MP2_CUDA(cudaMemcpyAsync(h_dst, d_src, bytes, cudaMemcpyDeviceToHost, s));
    return MP2_OK;
}

struct Streams {
    cudaStream_t copy = nullptr, comp = nullptr;
    cudaEvent_t ev = nullptr;
    int init() {
        MP2_CUDA(cudaStreamCreateWithFlags(&copy, cudaStreamNonBlocking));
        MP2_CUDA(cudaStreamCreateWithFlags(&comp, cudaStreamNonBlocking));
        MP2_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
        return MP2_OK;
    }
    ~Streams() {
        if (ev) cudaEventDestroy(ev);
        if (copy) cudaStreamDestroy(copy);
        if (comp) cudaStreamDestroy(comp);
    }
};

}  // namespace mp2

using namespace mp2;

extern "C" int mp2_tnf(const uint8_t *bases, const int64_t *offsets, int64_t n, int device,
                       uint32_t *counts, float *rel, float *gc) {
    if (n < 0) return fail(MP2_EINVAL, "negative n");
    if (!offsets) return fail(MP2_EINVAL, "offsets is NULL");
    if (n > INT32_MAX) return fail(MP2_EINVAL, "too many sequences in one call");
    const int64_t total = offsets[n] - offsets[0];
    if (offsets[0] != 0) return fail(MP2_EINVAL, "offsets[0] must be 0");
    for (int64_t i = 0; i < n; i++)
        if (offsets[i + 1] < offsets[i]) return fail(MP2_EINVAL, "offsets must be non-decreasing");
    if (total > 0 && !bases) return fail(MP2_EINVAL, "bases is NULL");
    MP2_TRY(require_device(device));
    if (n == 0) return MP2_OK;

    const bool trace = getenv("MP2_TRACE") != nullptr;
    const double t0 = omp_get_wtime();
    auto stamp = [&](const char *what) {
        if (trace) fprintf(stderr, "[mp2_tnf] %-28s %8.2f ms\n", what, (omp_get_wtime() - t0) * 1e3);
    };
    Streams st;
    MP2_TRY(st.init());
    stamp("streams");
    Temp d_bases, d_off, d_counts, d_gcc, d_rel, d_gc;
    MP2_TRY(d_bases.alloc((size_t)total + 16, st.copy));
    MP2_TRY(d_off.alloc(sizeof(int64_t) * (n + 1), st.copy));
    MP2_TRY(d_counts.alloc(sizeof(uint32_t) * MP2_NCANON * n, st.comp));
    MP2_TRY(d_gcc.alloc(sizeof(unsigned long long) * n, st.comp));
    if (rel) MP2_TRY(d_rel.alloc(sizeof(float) * MP2_NCANON * n, st.comp));
    if (gc) MP2_TRY(d_gc.alloc(sizeof(float) * n, st.comp));
    MP2_CUDA(cudaMemsetAsync(d_counts.p, 0, sizeof(uint32_t) * MP2_NCANON * n, st.comp));
    MP2_CUDA(cudaMemsetAsync(d_gcc.p, 0, sizeof(unsigned long long) * n, st.comp));

    stamp("alloc + memset enqueued");
    MP2_TRY(upload(d_off.p, offsets, sizeof(int64_t) * (n + 1), st.copy));
    const int64_t n_chunks = tnf_num_chunks(d_bases.p, total);
    int64_t next_chunk = 0;
    auto after = [&](size_t done) -> int {
        // chunks whose bytes have all been enqueued for copy can run behind that copy
        int64_t upto = ((int64_t)done == total) ? n_chunks : tnf_chunks_ready(d_bases.p, (int64_t)done, total);
        if (upto > next_chunk) {
            MP2_CUDA(cudaEventRecord(st.ev, st.copy));
            MP2_CUDA(cudaStreamWaitEvent(st.comp, st.ev, 0));
            MP2_TRY(tnf_launch_range(d_bases.as<uint8_t>(), d_off.as<int64_t>(), n, total,
                                     next_chunk, upto, d_counts.as<uint32_t>(),
                                     d_gcc.as<unsigned long long>(), st.comp, false));
            next_chunk = upto;
        }
        return MP2_OK;
    };
    MP2_TRY(upload(d_bases.p, bases, (size_t)total, st.copy, after));
    stamp("upload + kernels enqueued");
    if (trace) { cudaStreamSynchronize(st.copy); stamp("copy stream done"); cudaStreamSynchronize(st.comp); stamp("compute stream done"); }
    MP2_TRY(tnf_finalize(d_counts.as<uint32_t>(), d_gcc.as<unsigned long long>(),
                         d_off.as<int64_t>(), n, d_rel.as<float>(), d_gc.as<float>(), st.comp));
    if (counts) MP2_TRY(download(counts, d_counts.p, sizeof(uint32_t) * MP2_NCANON * n, st.comp));
    if (rel) MP2_TRY(download(rel, d_rel.p, sizeof(float) * MP2_NCANON * n, st.comp));
    if (gc) MP2_TRY(download(gc, d_gc.p, sizeof(float) * n, st.comp));
    MP2_CUDA(cudaStreamSynchronize(st.comp));
    MP2_CUDA(cudaStreamSynchronize(st.copy));
    stamp("results on host");
    return MP2_OK;
}

// Host-buffer form of the packed path.  What crosses PCIe: the codes (0.25 bytes per base) and either the
// validity mask (0.125 bytes per base) or, far smaller for real genomes, the runs of invalid bases from which
// the device rebuilds the mask.  The codes are uploaded in slices of 64 Mbases; the pieces wholly inside what has
// been enqueued run behind the copy of the next slice, and the rows of the contigs that are complete by then are
// finalised and copied back on a third stream, so the device-to-host copy of the results runs under the upload.
extern "C" int mp2_tnf_packed(const uint32_t *codes, const uint32_t *valid, const int64_t *invalid_runs, int64_t n_runs,
                              const int64_t *offsets, int64_t n, int device, uint32_t *counts, float *rel, float *gc) {
    if (n < 0 || n_runs < 0) return fail(MP2_EINVAL, "negative size");
    if (!offsets) return fail(MP2_EINVAL, "offsets is NULL");
    if (n > INT32_MAX) return fail(MP2_EINVAL, "too many sequences in one call");
    if (offsets[0] != 0) return fail(MP2_EINVAL, "offsets[0] must be 0");
    for (int64_t i = 0; i < n; i++)
        if (offsets[i + 1] < offsets[i]) return fail(MP2_EINVAL, "offsets must be non-decreasing");
    const int64_t total = offsets[n];
    if (total > 0 && !codes) return fail(MP2_EINVAL, "codes is NULL");
    if (total > 0 && !valid && !(invalid_runs || n_runs == 0)) return fail(MP2_EINVAL, "neither valid nor invalid_runs given");
    if (!valid)
        for (int64_t r = 0; r < n_runs; r++)
            if (invalid_runs[2 * r] < 0 || invalid_runs[2 * r + 1] > total || invalid_runs[2 * r + 1] < invalid_runs[2 * r])
                return fail(MP2_EINVAL, "invalid_runs[%lld] is not inside the stream", (long long)r);
    MP2_TRY(require_device(device));
    if (n == 0) return MP2_OK;

    Streams st;
    MP2_TRY(st.init());
    cudaStream_t out = nullptr;
    MP2_CUDA(cudaStreamCreateWithFlags(&out, cudaStreamNonBlocking));
    struct OutGuard { cudaStream_t s; ~OutGuard() { if (s) cudaStreamDestroy(s); } } og{out};
    cudaEvent_t rows_done = nullptr;
    MP2_CUDA(cudaEventCreateWithFlags(&rows_done, cudaEventDisableTiming));
    struct EvGuard { cudaEvent_t e; ~EvGuard() { if (e) cudaEventDestroy(e); } } eg{rows_done};

    const int64_t nq = (total + 63) / 64;
    Temp d_codes, d_valid, d_runs, d_off, d_counts, d_gcc, d_rel, d_gc;
    MP2_TRY(d_codes.alloc(sizeof(uint32_t) * 4 * (size_t)nq + 16, st.copy));
    MP2_TRY(d_valid.alloc(sizeof(uint32_t) * 2 * (size_t)nq + 16, st.copy));
    MP2_TRY(d_off.alloc(sizeof(int64_t) * (n + 1), st.copy));
    MP2_TRY(d_counts.alloc(sizeof(uint32_t) * MP2_NCANON * n, st.comp));
    MP2_TRY(d_gcc.alloc(sizeof(unsigned long long) * n, st.comp));
    if (rel) MP2_TRY(d_rel.alloc(sizeof(float) * MP2_NCANON * n, st.comp));
    if (gc) MP2_TRY(d_gc.alloc(sizeof(float) * n, st.comp));
    MP2_CUDA(cudaMemsetAsync(d_counts.p, 0, sizeof(uint32_t) * MP2_NCANON * n, st.comp));
    MP2_CUDA(cudaMemsetAsync(d_gcc.p, 0, sizeof(unsigned long long) * n, st.comp));
    MP2_TRY(upload(d_off.p, offsets, sizeof(int64_t) * (n + 1), st.copy));
    if (valid) {
        MP2_TRY(upload(d_valid.p, valid, sizeof(uint32_t) * 2 * (size_t)nq, st.copy));
    } else {  // the whole mask before the first slice: a memset and a few thousand word updates
        int64_t tail[2] = {total, nq * 64};
        MP2_TRY(d_runs.alloc(sizeof(int64_t) * 2 * (size_t)(n_runs + 1), st.copy));
        if (n_runs) MP2_TRY(upload(d_runs.p, invalid_runs, sizeof(int64_t) * 2 * (size_t)n_runs, st.copy));
        MP2_CUDA(cudaMemcpyAsync(d_runs.as<int64_t>() + 2 * n_runs, tail, sizeof tail, cudaMemcpyHostToDevice, st.copy));
        MP2_CUDA(cudaStreamSynchronize(st.copy));  // `tail` lives on this stack frame
        MP2_TRY(tnf_valid_from_runs(d_runs.as<int64_t>(), n_runs + 1, total, d_valid.as<uint32_t>(), st.copy));
    }
    const int64_t n_chunks = tnf_packed_num_chunks(total);
    const int64_t QS = 1 << 20;  // quads per slice (64 Mbases: 16 MB of codes)
    int64_t next_chunk = 0, rows_out = 0;
    // rows go home early only into page-locked outputs: a copy into pageable memory blocks this thread until the
    // slice's kernels are done, which would serialise the upload behind the compute
    const bool early_out = (!counts || is_pinned(counts)) && (!rel || is_pinned(rel)) && (!gc || is_pinned(gc));
    const bool trace = getenv("MP2_TRACE") != nullptr;
    const double t0 = omp_get_wtime();
    auto stamp = [&](const char *what) {
        if (trace) fprintf(stderr, "[mp2_tnf_packed] %-28s %8.2f ms\n", what, (omp_get_wtime() - t0) * 1e3);
    };
    for (int64_t q0 = 0; q0 < nq; q0 += QS) {
        const int64_t q1 = std::min(nq, q0 + QS);
        MP2_TRY(upload(d_codes.as<uint32_t>() + 4 * q0, codes + 4 * q0, sizeof(uint32_t) * 4 * (size_t)(q1 - q0), st.copy));
        const bool last = q1 == nq;
        const int64_t upto = last ? n_chunks : tnf_packed_chunks_ready(q1 * 64, total);
        if (upto > next_chunk) {
            MP2_CUDA(cudaEventRecord(st.ev, st.copy));
            MP2_CUDA(cudaStreamWaitEvent(st.comp, st.ev, 0));
            MP2_TRY(tnf_packed_launch_range(d_codes.as<uint32_t>(), d_valid.as<uint32_t>(), d_off.as<int64_t>(), n, total,
                                            next_chunk, upto, d_counts.as<uint32_t>(), d_gcc.as<unsigned long long>(),
                                            st.comp));
            next_chunk = upto;
        }
        // rows whose contig ends inside the pieces counted so far are final: normalise and send them home now
        const int64_t done_bases = tnf_packed_bases_done(next_chunk, total);
        int64_t rows_now = last ? n : early_out ? (int64_t)(std::upper_bound(offsets + 1, offsets + n + 1, done_bases) - (offsets + 1)) : 0;
        if (rows_now > n) rows_now = n;
        if (rows_now > rows_out) {
            const int64_t r0 = rows_out, k = rows_now - rows_out;
            MP2_TRY(tnf_finalize(d_counts.as<uint32_t>() + r0 * MP2_NCANON, d_gcc.as<unsigned long long>() + r0,
                                 d_off.as<int64_t>() + r0, k, rel ? d_rel.as<float>() + r0 * MP2_NCANON : nullptr,
                                 gc ? d_gc.as<float>() + r0 : nullptr, st.comp));
            MP2_CUDA(cudaEventRecord(rows_done, st.comp));
            MP2_CUDA(cudaStreamWaitEvent(out, rows_done, 0));
            if (counts) MP2_TRY(download(counts + r0 * MP2_NCANON, d_counts.as<uint32_t>() + r0 * MP2_NCANON, sizeof(uint32_t) * MP2_NCANON * k, out));
            if (rel) MP2_TRY(download(rel + r0 * MP2_NCANON, d_rel.as<float>() + r0 * MP2_NCANON, sizeof(float) * MP2_NCANON * k, out));
            if (gc) MP2_TRY(download(gc + r0, d_gc.as<float>() + r0, sizeof(float) * k, out));
            rows_out = rows_now;
        }
    }
    stamp("everything enqueued");
    if (trace) { cudaStreamSynchronize(st.copy); stamp("copy stream done"); cudaStreamSynchronize(st.comp); stamp("compute stream done"); }
    MP2_CUDA(cudaStreamSynchronize(st.comp));
    MP2_CUDA(cudaStreamSynchronize(out));
    MP2_CUDA(cudaStreamSynchronize(st.copy));
    stamp("results on host");
    return MP2_OK;
}

extern "C" int mp2_cov_events(const uint32_t *starts, const uint32_t *ends, const int64_t *ev_off,
                              const int64_t *contig_off, int64_t n_contigs, int64_t n_events,
                              uint32_t max_span, uint32_t max_disorder, uint32_t end_excl, float trim_lower,
                              float trim_upper, int device, float *cov, int64_t cov_stride, mp2_cov_stats *stats) {
    if (n_contigs < 0 || n_events < 0) return fail(MP2_EINVAL, "negative size");
    if (!contig_off || !ev_off || !cov) return fail(MP2_EINVAL, "NULL argument");
    if (n_events > 0 && (!starts || !ends)) return fail(MP2_EINVAL, "NULL events");
    if (cov_stride < 1) return fail(MP2_EINVAL, "cov_stride must be >= 1");
    if (contig_off[0] != 0 || ev_off[0] != 0) return fail(MP2_EINVAL, "offsets must start at 0");
    if (ev_off[n_contigs] != n_events) return fail(MP2_EINVAL, "ev_off[n_contigs] != n_events");
    for (int64_t i = 0; i < n_contigs; i++)
        if (contig_off[i + 1] < contig_off[i] || ev_off[i + 1] < ev_off[i])
            return fail(MP2_EINVAL, "offsets must be non-decreasing");
    MP2_TRY(require_device(device));
    if (n_contigs == 0) return MP2_OK;
    Streams st;
    MP2_TRY(st.init());
    Temp d_s, d_e, d_eo, d_co, d_cov, d_stats;
    MP2_TRY(d_s.alloc(sizeof(uint32_t) * n_events, st.comp));
    MP2_TRY(d_e.alloc(sizeof(uint32_t) * n_events, st.comp));
    MP2_TRY(d_eo.alloc(sizeof(int64_t) * (n_contigs + 1), st.comp));
    MP2_TRY(d_co.alloc(sizeof(int64_t) * (n_contigs + 1), st.comp));
    MP2_TRY(d_cov.alloc(sizeof(float) * n_contigs, st.comp));
    if (stats) MP2_TRY(d_stats.alloc(sizeof(mp2_cov_stats) * n_contigs, st.comp));
    MP2_TRY(upload(d_eo.p, ev_off, sizeof(int64_t) * (n_contigs + 1), st.comp));
    MP2_TRY(upload(d_co.p, contig_off, sizeof(int64_t) * (n_contigs + 1), st.comp));
    MP2_TRY(upload(d_s.p, starts, sizeof(uint32_t) * n_events, st.comp));
    MP2_TRY(upload(d_e.p, ends, sizeof(uint32_t) * n_events, st.comp));
    MP2_TRY(mp2_cov_events_device(d_s.p, d_e.p, d_eo.p, d_co.p, n_contigs, n_events, contig_off[n_contigs],
                                  max_span, max_disorder, end_excl, trim_lower, trim_upper, d_cov.p, 1, d_stats.p,
                                  st.comp));
    MP2_CUDA(cudaStreamSynchronize(st.comp));
    if (cov_stride == 1) {
        MP2_TRY(download(cov, d_cov.p, sizeof(float) * n_contigs, st.comp));
    } else {
        MP2_CUDA(cudaMemcpy2DAsync(cov, sizeof(float) * cov_stride, d_cov.p, sizeof(float), sizeof(float),
                                   n_contigs, cudaMemcpyDeviceToHost, st.comp));
    }
    if (stats) MP2_TRY(download(stats, d_stats.p, sizeof(mp2_cov_stats) * n_contigs, st.comp));
    MP2_CUDA(cudaStreamSynchronize(st.comp));
    return MP2_OK;
}

extern "C" int mp2_logratio(const float *data, const int64_t *weights, const int64_t *mag_off,
                            int64_t n_mags, int64_t n_rows, int64_t n_cols, const uint8_t *col_mask,
                            double base, float pseudo, double clip_lo, double clip_hi, int64_t min_rows,
                            int one_sided, int device, double *scores, float *medians) {
    if (n_mags < 0 || n_rows < 0 || n_cols < 1) return fail(MP2_EINVAL, "bad shape");
    if (!mag_off || !scores || (n_rows > 0 && (!data || !weights))) return fail(MP2_EINVAL, "NULL argument");
    if (mag_off[0] != 0 || mag_off[n_mags] != n_rows) return fail(MP2_EINVAL, "mag_off must span [0, n_rows]");
    for (int64_t i = 0; i < n_mags; i++)
        if (mag_off[i + 1] < mag_off[i]) return fail(MP2_EINVAL, "mag_off must be non-decreasing");
    MP2_TRY(require_device(device));
    if (n_rows == 0) return MP2_OK;
    Streams st;
    MP2_TRY(st.init());
    Temp d_data, d_w, d_mo, d_mask, d_sc, d_med;
    MP2_TRY(d_data.alloc(sizeof(float) * n_rows * n_cols, st.comp));
    MP2_TRY(d_w.alloc(sizeof(int64_t) * n_rows, st.comp));
    MP2_TRY(d_mo.alloc(sizeof(int64_t) * (n_mags + 1), st.comp));
    MP2_TRY(d_sc.alloc(sizeof(double) * n_rows, st.comp));
    MP2_TRY(d_med.alloc(sizeof(float) * n_mags * n_cols, st.comp));
    if (col_mask) {
        MP2_TRY(d_mask.alloc(n_mags * n_cols, st.comp));
        MP2_TRY(upload(d_mask.p, col_mask, n_mags * n_cols, st.comp));
    }
    MP2_TRY(upload(d_data.p, data, sizeof(float) * n_rows * n_cols, st.comp));
    MP2_TRY(upload(d_w.p, weights, sizeof(int64_t) * n_rows, st.comp));
    MP2_TRY(upload(d_mo.p, mag_off, sizeof(int64_t) * (n_mags + 1), st.comp));
    MP2_TRY(mp2_logratio_device(d_data.p, d_w.p, d_mo.p, n_mags, n_rows, n_cols, col_mask ? d_mask.p : nullptr,
                                base, pseudo, clip_lo, clip_hi, min_rows, one_sided, d_sc.p, d_med.p, st.comp));
    MP2_TRY(download(scores, d_sc.p, sizeof(double) * n_rows, st.comp));
    if (medians) MP2_TRY(download(medians, d_med.p, sizeof(float) * n_mags * n_cols, st.comp));
    MP2_CUDA(cudaStreamSynchronize(st.comp));
    return MP2_OK;
}
