// Micro-benchmark: random 32-byte gathers (LDG.256) over a table of S bytes, all SMs.  Answers: gather rate when the table is
// L2-resident, rate from DRAM, and the effective L2 capacity for a table every SM reads.  Build: nvcc -arch=sm_100a -O3.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
typedef unsigned long long u64;
struct u64x4 { u64 a, b, c, d; };
__device__ __forceinline__ u64x4 ld_v4(const u64 *p) {
    u64x4 v;
    asm volatile("ld.global.nc.v4.u64 {%0, %1, %2, %3}, [%4];" : "=l"(v.a), "=l"(v.b), "=l"(v.c), "=l"(v.d) : "l"(p));
    return v;
}
__device__ __forceinline__ u64x4 ld_v4_hint(const u64 *p, u64 pol) {
    u64x4 v;
    asm volatile("ld.global.L2::cache_hint.v4.u64 {%0, %1, %2, %3}, [%4], %5;" : "=l"(v.a), "=l"(v.b), "=l"(v.c), "=l"(v.d) : "l"(p), "l"(pol));
    return v;
}
__device__ __forceinline__ u64 mix(u64 x) { x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33; return x; }
template <int U, bool HINT>
__global__ void __launch_bounds__(256) k_gather(const u64 *__restrict__ T, u64 nrows, u64 iters, u64 *__restrict__ out, u64 seed) {
    u64 pol = 0;
    if (HINT) asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    u64 t = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    u64 acc = 0;
    for (u64 i = 0; i < iters; i++) {
        u64x4 v[U];
#pragma unroll
        for (int u = 0; u < U; u++) {
            u64 r = __umul64hi(mix(seed + (t * iters + i) * U + u), nrows);
            v[u] = HINT ? ld_v4_hint(T + r * 4, pol) : ld_v4(T + r * 4);
        }
#pragma unroll
        for (int u = 0; u < U; u++) acc |= v[u].a | v[u].b | v[u].c | v[u].d;
    }
    if (acc == 0x123456789ULL) out[0] = acc;
}
int main() {
    cudaSetDevice(0);
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    printf("SMs %d L2 %.1f MB persistingL2CacheMax %.1f MB\n", p.multiProcessorCount, p.l2CacheSize / 1048576.0, p.persistingL2CacheMaxSize / 1048576.0);
    const size_t MAXB = (size_t)512 << 20;
    u64 *T, *out; cudaMalloc(&T, MAXB); cudaMalloc(&out, 8); cudaMemset(T, 0, MAXB);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    size_t sizes[] = {4, 8, 16, 24, 32, 48, 64, 80, 96, 128, 192, 270, 512};
    for (int hint = 0; hint < 2; hint++)
    for (int occ = 0; occ < 2; occ++)
    for (size_t s : sizes) {
        u64 nrows = (s << 20) / 32;
        int grid = p.multiProcessorCount * (occ ? 8 : 4);
        u64 iters = 256;
        auto run = [&](u64 seed) { if (hint) k_gather<4, true><<<grid, 256>>>(T, nrows, iters, out, seed); else k_gather<4, false><<<grid, 256>>>(T, nrows, iters, out, seed); };
        run(1); run(2);
        cudaEventRecord(e0);
        for (int r = 0; r < 3; r++) run(3 + r);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        double g = 3.0 * grid * 256.0 * iters * 4;
        printf("hint %d ctas/SM %d table %4zu MB : %7.1f Ggather/s  (%6.2f TB/s)\n", hint, occ ? 8 : 4, s, g / ms / 1e6, g * 32 / ms / 1e9);
    }
    return 0;
}
