// Micro-benchmark 2: skewed random 32-byte gathers (fraction `phot` of them into the first `hot` MB of a 270 MB table) with a
// concurrent once-only stream (one 32-byte read per 4 gathers, the col_idx/Y traffic of the pull kernel).  Which L2 eviction
// controls keep the hot prefix resident?  gather modes: 0 plain, 1 .L2::evict_last on every gather, 2 evict_last hot /
// evict_first cold; stream modes: 0 plain, 1 .L1::no_allocate.L2::evict_first.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
typedef unsigned long long u64;
typedef unsigned int u32;
struct u64x4 { u64 a, b, c, d; };
#define LD4(QUAL, v, p) asm volatile("ld.global" QUAL ".v4.u64 {%0, %1, %2, %3}, [%4];" : "=l"(v.a), "=l"(v.b), "=l"(v.c), "=l"(v.d) : "l"(p))
__device__ __forceinline__ u64 mix(u64 x) { x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33; return x; }
template <int GM, int SM>
__global__ void __launch_bounds__(256) k(const u64 *__restrict__ T, u64 nhot, u64 nrows, u32 phot1024, const u64 *__restrict__ S, u64 srows,
                                         u64 iters, u64 *__restrict__ out, u64 seed) {
    u64 t = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    u64 nthreads = (u64)gridDim.x * blockDim.x;
    u64 acc = 0;
    for (u64 i = 0; i < iters; i++) {
        u64x4 v[4], s;
        u64 si = i * nthreads + t; if (si >= srows) si -= srows * (si / srows);
        const u64 *sp = S + si * 4;
        if (SM == 0) LD4(".nc", s, sp); else LD4(".L1::no_allocate.L2::evict_first", s, sp);
#pragma unroll
        for (int u = 0; u < 4; u++) {
            u64 h = mix(seed + (t * iters + i) * 4 + u);
            bool hot = (mix(h) & 1023) < phot1024;
            u64 r = hot ? __umul64hi(h, nhot) : nhot + __umul64hi(h, nrows - nhot);   // h is uniform in 2^64
            const u64 *p = T + r * 4;
            if (GM == 0) LD4(".nc", v[u], p);
            else if (GM == 1) LD4(".L2::evict_last", v[u], p);
            else { if (hot) LD4(".L2::evict_last", v[u], p); else LD4(".L2::evict_first", v[u], p); }
        }
        acc |= s.a | s.b | s.c | s.d;
#pragma unroll
        for (int u = 0; u < 4; u++) acc |= v[u].a | v[u].b | v[u].c | v[u].d;
    }
    if (acc == 0x123456789ULL) out[0] = acc;
}
typedef void (*KF)(const u64 *, u64, u64, u32, const u64 *, u64, u64, u64 *, u64);
int main() {
    cudaSetDevice(0);
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    const size_t TB = (size_t)270 << 20, SB = (size_t)2 << 30;
    u64 *T, *S, *out; cudaMalloc(&T, TB); cudaMalloc(&S, SB); cudaMalloc(&out, 8); cudaMemset(T, 0, TB); cudaMemset(S, 0, SB);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    KF ks[3][2] = {{k<0, 0>, k<0, 1>}, {k<1, 0>, k<1, 1>}, {k<2, 0>, k<2, 1>}};
    int hots[] = {16, 32, 48, 64};
    for (int hm : hots)
        for (int gm = 0; gm < 3; gm++)
            for (int sm = 0; sm < 2; sm++) {
                u64 nrows = TB / 32, nhot = ((size_t)hm << 20) / 32, srows = SB / 32;
                // edge share of the hottest hm MB of RMAT-24 at 32 B per vertex: 16 MB 78%, 32 MB 87%, 48 MB 91%, 64 MB 94%
                u32 ph = hm == 16 ? 800 : hm == 32 ? 890 : hm == 48 ? 930 : 960;
                int grid = p.multiProcessorCount * 5;
                u64 iters = 256;
                ks[gm][sm]<<<grid, 256>>>(T, nhot, nrows, ph, S, srows, iters, out, 1);
                cudaEventRecord(e0);
                for (int r = 0; r < 3; r++) ks[gm][sm]<<<grid, 256>>>(T, nhot, nrows, ph, S, srows, iters, out, 2 + r);
                cudaEventRecord(e1); cudaEventSynchronize(e1);
                float ms; cudaEventElapsedTime(&ms, e0, e1);
                double g = 3.0 * grid * 256.0 * iters * 4;
                printf("hot %2d MB (%.0f%% of gathers) gather-mode %d stream-mode %d : %7.1f Ggather/s\n", hm, ph / 10.24, gm, sm, g / ms / 1e6);
            }
    return 0;
}
