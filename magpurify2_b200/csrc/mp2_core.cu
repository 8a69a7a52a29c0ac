// libmp2 core: error state, device selection, stream-ordered temporaries, canonical tables.
#include <atomic>
#include <map>
#include <mutex>
#include <vector>

#include "mp2_common.h"

namespace mp2 {

static thread_local std::string g_err;

void set_error(const char *fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
}

int fail(int code, const char *fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}

int require_device(int device) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n <= 0) {
        (void)cudaGetLastError();
        return fail(MP2_ENODEV,
                    "no usable CUDA device (%s); libmp2 has no CPU fallback",
                    e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
    }
    if (device >= 0) {
        if (device >= n) return fail(MP2_EINVAL, "device %d out of range (have %d)", device, n);
        MP2_CUDA(cudaSetDevice(device));
    }
    // keep stream-ordered temporaries cached between calls
    static std::once_flag once[64];
    int cur = 0;
    MP2_CUDA(cudaGetDevice(&cur));
    if (cur < 64) {
        std::call_once(once[cur], [cur]() {
            cudaMemPool_t pool;
            if (cudaDeviceGetDefaultMemPool(&pool, cur) == cudaSuccess) {
                uint64_t thr = UINT64_MAX;
                cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
            }
            (void)cudaGetLastError();
        });
    }
    return MP2_OK;
}

int num_sms() {
    static int cached[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    if (dev < 64 && cached[dev]) return cached[dev];
    int n = 148;
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (dev < 64) cached[dev] = n;
    return n;
}

int Temp::alloc(size_t bytes, cudaStream_t stream) {
    s = stream;
    if (bytes == 0) bytes = 16;
    MP2_CUDA(cudaMallocAsync(&p, bytes, stream));
    return MP2_OK;
}

Temp::~Temp() {
    if (p) cudaFreeAsync(p, s);
}

// ---- launch bookkeeping / optional per-kernel timing --------------------------------------
struct ProfRec {
    const char *name;
    cudaEvent_t a, b;
    int dev;
};
static std::mutex g_prof_mu;
static bool g_prof_on = false;
static std::vector<ProfRec> g_prof_pending;
static std::vector<cudaEvent_t> g_prof_free;
static std::map<std::string, std::pair<double, int64_t>> g_prof_acc;
static std::atomic<int64_t> g_launches{0};

static cudaEvent_t prof_event() {
    if (!g_prof_free.empty()) {
        cudaEvent_t e = g_prof_free.back();
        g_prof_free.pop_back();
        return e;
    }
    cudaEvent_t e = nullptr;
    cudaEventCreate(&e);
    return e;
}

Launch::Launch(const char *name_, cudaStream_t stream) : name(name_), s(stream), slot(nullptr) {
    g_launches.fetch_add(1, std::memory_order_relaxed);
    if (!g_prof_on) return;
    std::lock_guard<std::mutex> lk(g_prof_mu);
    ProfRec *r = new ProfRec{name, prof_event(), prof_event(), 0};
    cudaGetDevice(&r->dev);
    cudaEventRecord(r->a, s);
    slot = r;
}

Launch::~Launch() {
    if (!slot) return;
    ProfRec *r = (ProfRec *)slot;
    cudaEventRecord(r->b, s);
    std::lock_guard<std::mutex> lk(g_prof_mu);
    g_prof_pending.push_back(*r);
    delete r;
}

static void prof_drain() {
    for (auto &r : g_prof_pending) {
        float ms = 0.f;
        if (cudaEventSynchronize(r.b) == cudaSuccess && cudaEventElapsedTime(&ms, r.a, r.b) == cudaSuccess) {
            auto &acc = g_prof_acc[r.name];
            acc.first += ms;
            acc.second += 1;
        }
        (void)cudaGetLastError();
        g_prof_free.push_back(r.a);
        g_prof_free.push_back(r.b);
    }
    g_prof_pending.clear();
}

// ---- canonical 4-mer tables, generated from the rank rule -------------------------------
// code: 2 bits per base, A=0 C=1 G=2 T=3, first base most significant.
// class(code) = rank of min(code, revcomp(code)) among canonical codes in ascending order,
// which is what REV_COMPLEMENTARY_FOURMERS (src/composition.rs:24-37) tabulates.
static uint8_t g_lut[256];
static uint8_t g_canon[MP2_NCANON];
static std::once_flag g_tab_once;

static unsigned revcomp4(unsigned c) {
    unsigned r = 0;
    for (int i = 0; i < 4; i++) {
        r = (r << 2) | (3u - (c & 3u));
        c >>= 2;
    }
    return r;
}

static void build_tables() {
    int next = 0;
    int rank[256];
    for (unsigned c = 0; c < 256; c++) {
        unsigned rc = revcomp4(c);
        if (c <= rc) {
            g_canon[next] = (uint8_t)c;
            rank[c] = next++;
        } else {
            rank[c] = -1;
        }
    }
    for (unsigned c = 0; c < 256; c++) {
        unsigned rc = revcomp4(c);
        g_lut[c] = (uint8_t)rank[c <= rc ? c : rc];
    }
}

const uint8_t *host_lut256() {
    std::call_once(g_tab_once, build_tables);
    return g_lut;
}
const uint8_t *host_canon_code136() {
    std::call_once(g_tab_once, build_tables);
    return g_canon;
}

}  // namespace mp2

extern "C" {

const char *mp2_last_error(void) { return mp2::g_err.c_str(); }

const char *mp2_version(void) { return "magpurify2-b200 0.1.0 (sm_100a)"; }

int mp2_device_count(void) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) {
        (void)cudaGetLastError();
        return mp2::fail(MP2_ENODEV, "cudaGetDeviceCount: %s", cudaGetErrorString(e));
    }
    return n;
}

int mp2_profile_enable(int on) {
    std::lock_guard<std::mutex> lk(mp2::g_prof_mu);
    mp2::g_prof_on = on != 0;
    return MP2_OK;
}

int mp2_profile_reset(void) {
    std::lock_guard<std::mutex> lk(mp2::g_prof_mu);
    mp2::prof_drain();
    mp2::g_prof_acc.clear();
    return MP2_OK;
}

int mp2_profile_read(const char *prefix, double *ms, int64_t *launches) {
    std::lock_guard<std::mutex> lk(mp2::g_prof_mu);
    mp2::prof_drain();
    std::string p = prefix ? prefix : "";
    double t = 0;
    int64_t n = 0;
    for (auto &kv : mp2::g_prof_acc)
        if (kv.first.compare(0, p.size(), p) == 0) {
            t += kv.second.first;
            n += kv.second.second;
        }
    if (ms) *ms = t;
    if (launches) *launches = n;
    return MP2_OK;
}

int64_t mp2_launch_count(void) { return mp2::g_launches.load(); }

int mp2_canonical_lut(uint8_t lut[256]) {
    if (!lut) return mp2::fail(MP2_EINVAL, "lut is NULL");
    const uint8_t *t = mp2::host_lut256();
    for (int i = 0; i < 256; i++) lut[i] = t[i];
    return MP2_OK;
}

}  // extern "C"
