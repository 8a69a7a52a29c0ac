// Micro-benchmark: shared-memory histogram update rates on sm_100a (design input for the
// tnf kernel).  Not part of the product.  Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

__device__ __forceinline__ uint32_t rnd(uint32_t &x) { x = x * 1664525u + 1013904223u; return x >> 24; }

template <int MODE>
__global__ void __launch_bounds__(256) k(int iters, uint32_t *out) {
    extern __shared__ uint32_t sm[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int words = (MODE == 10 || MODE == 13) ? 8192 : (MODE == 11) ? 32768 : (MODE == 12) ? 4096 : (MODE == 4) ? 256 : (MODE == 2) ? 256 * 8 : (MODE == 1 ? 128 * 32 : (MODE == 5 ? 64 * 32 : (MODE == 3 ? 2 * 256 * 32 : 256 * 32)));
    for (int i = tid; i < words; i += 256) sm[i] = 0;
    __syncthreads();
    uint32_t x = tid * 2654435761u + blockIdx.x * 40503u + 12345u;
    uint32_t acc = 0;
    uint8_t *smb = reinterpret_cast<uint8_t *>(sm);
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < 16; u++) {
            uint32_t c = rnd(x);
            if (MODE == 0) atomicAdd(&sm[c * 32 + lane], 1u);                       // conflict-free u32
            else if (MODE == 1) atomicAdd(&sm[(c >> 1) * 32 + lane], 1u << (16 * (c & 1)));  // packed u16
            else if (MODE == 2) atomicAdd(&sm[warp * 256 + c], 1u);                 // per-warp compact
            else if (MODE == 3) {                                                  // private bytes, no atomics
                uint8_t *p = smb + ((c * 32 + lane) * 4 + (warp & 3)) + (warp >> 2) * 32768;
                *p = *p + 1;
            } else if (MODE == 4) acc ^= c;                                        // ALU only
            else if (MODE == 5) atomicAdd(&sm[(c >> 2) * 32 + lane], 1u << (8 * (c & 3)));   // packed u8
            else if (MODE == 6) {                                                  // non-atomic u32 RMW (racy; rate only)
                uint32_t *p = &sm[c * 32 + lane];
                *p = *p + 1;
            } else if (MODE == 8) atomicAdd(&sm[(c & 7)], 1u);                      // 8 hot addresses, no privatisation
            else if (MODE == 9) atomicAdd(&sm[(c & 15) * 32 + lane], 1u);            // few rows, lane-private
            else if (MODE == 10) { uint32_t c2 = (c << 2) | (x >> 22 & 3); atomicAdd(&sm[warp * 1024 + c2], 1u); }   // per-warp [1024]
            else if (MODE == 11) { uint32_t c2 = (c << 4) | (x >> 20 & 15); atomicAdd(&sm[warp * 4096 + c2], 1u); }  // per-warp [4096]
            else if (MODE == 12) atomicAdd(&sm[warp * 512 + c * 2 + (lane & 1)], 1u);    // per-warp [256] x2 replicas
            else if (MODE == 13) atomicAdd(&sm[warp * 1024 + c * 4 + (lane & 3)], 1u);   // per-warp [256] x4 replicas
            else if (MODE == 14) atomicAdd(&sm[c], 1u);                                  // one [256] per CTA
            else if (MODE == 15) { uint32_t cc = (c & (c >> 1)) ; atomicAdd(&sm[warp * 256 + cc], 1u); }  // skewed codes
            else if (MODE == 7) atomicAdd(&sm[c * 32 + ((lane + c) & 31)], 1u);     // conflict-free rotated
        }
        if (MODE == 3) asm volatile("" ::: "memory");
    }
    __syncthreads();
    uint32_t s = acc;
    for (int i = tid; i < words; i += 256) s += sm[i];
    if (s == 0xdeadbeef) out[0] = s;
}

template <int MODE>
void run(const char *name, int smem, int blocks_per_sm, int iters) {
    int dev = 0, sms = 0, khz = 0;
    CK(cudaGetDevice(&dev));
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    CK(cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev));
    CK(cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    int occ = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k<MODE>, 256, smem));
    if (blocks_per_sm > occ) blocks_per_sm = occ;
    uint32_t *out; CK(cudaMalloc(&out, 4));
    cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    int grid = sms * blocks_per_sm;
    k<MODE><<<grid, 256, smem>>>(iters / 4, out);
    CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < 3; r++) {
        CK(cudaEventRecord(a));
        k<MODE><<<grid, 256, smem>>>(iters, out);
        CK(cudaEventRecord(b));
        CK(cudaEventSynchronize(b));
        float ms; CK(cudaEventElapsedTime(&ms, a, b));
        if (ms < best) best = ms;
    }
    double upd = (double)grid * 256 * iters * 16;
    double gups = upd / (best * 1e-3) / 1e9;
    printf("%-28s occ=%d ctas/sm=%d  %8.3f ms  %8.1f Gupd/s  %6.2f upd/clk/SM@%dMHz  (=%.0f%% of 6551 GB/s at 1 B/base)\n",
           name, occ, blocks_per_sm, best, gups, gups * 1e9 / sms / (khz * 1e3), khz / 1000, gups / 6551 * 100);
    CK(cudaFree(out));
}

int main() {
    int iters = 2000;
    for (int bps = 2; bps <= 8; bps *= 4) {
        printf("--- %d CTA(s)/SM x 256 threads\n", bps);
        run<4>("alu_only", 1024, bps, iters);
        run<0>("atoms_u32_[code][lane]", 32768, bps, iters);
        run<7>("atoms_u32_rotated", 32768, bps, iters);
        run<1>("atoms_packed16_[128][lane]", 16384, bps, iters);
        run<5>("atoms_packed8_[64][lane]", 8192, bps, iters);
        run<2>("atoms_per_warp_[256]", 8192, bps, iters);
        run<3>("rmw_private_u8", 65536, bps, iters);
        run<6>("rmw_nonatomic_u32", 32768, bps, iters);
        run<8>("atoms_u32_same_addr_heavy", 32768, bps, iters);
        run<9>("atoms_u32_[16][lane]", 32768, bps, iters);
        run<10>("atoms_per_warp_[1024]", 32768, bps, iters);
        run<11>("atoms_per_warp_[4096]", 131072, bps, iters);
        run<12>("atoms_per_warp_[256]x2", 16384, bps, iters);
        run<13>("atoms_per_warp_[256]x4", 32768, bps, iters);
        run<14>("atoms_per_cta_[256]", 8192, bps, iters);
        run<15>("atoms_per_warp_[256]_skewed", 8192, bps, iters);
    }
    return 0;
}
