// Internal helpers shared by the libmp2 translation units (not part of the C ABI).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdarg>
#include <cstdio>
#include <string>

#include "mp2.h"

namespace mp2 {

void set_error(const char *fmt, ...);
int fail(int code, const char *fmt, ...);

#define MP2_CUDA(expr)                                                                        \
    do {                                                                                      \
        cudaError_t _e = (expr);                                                              \
        if (_e != cudaSuccess) {                                                              \
            (void)cudaGetLastError();                                                         \
            return ::mp2::fail(_e == cudaErrorMemoryAllocation ? MP2_ENOMEM : MP2_ECUDA,      \
                               "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e),        \
                               __FILE__, __LINE__);                                           \
        }                                                                                     \
    } while (0)

#define MP2_TRY(expr)            \
    do {                         \
        int _r = (expr);         \
        if (_r != MP2_OK) return _r; \
    } while (0)

// Makes sure a device is usable and selected; the product has no CPU path.
int require_device(int device);
int num_sms();

// Stream-ordered temporary (cudaMallocAsync on the call's stream, freed in the destructor).
struct Temp {
    void *p = nullptr;
    cudaStream_t s = nullptr;
    int alloc(size_t bytes, cudaStream_t stream);
    ~Temp();
    template <class T> T *as() const { return reinterpret_cast<T *>(p); }
};

// Kernel launch bookkeeping (mp2_launch_count / mp2_profile_*).  Usage:
//   { Launch L("tnf_kernel", stream); kernel<<<g, b, 0, stream>>>(...); }
struct Launch {
    const char *name;
    cudaStream_t s;
    void *slot;
    Launch(const char *name, cudaStream_t stream);
    ~Launch();
};

// canonical tables (host copies; device copies live in __constant__ of the TU that uses them)
const uint8_t *host_lut256();         // code -> class
const uint8_t *host_canon_code136();  // class -> smallest code of the class

}  // namespace mp2
