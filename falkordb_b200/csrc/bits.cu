// bits.cu -- frontier bit-matrix path.
//
// FalkorDB's CondTraverse drives GrB_mxm with a short-fat left operand: F is |batch| x n with
// |batch| <= 1024 rows (graph/src/runtime/batch.rs:81, runtime/ops/cond_traverse.rs:600-605).
// For such operands the row-wise product F*A is the multi-source frontier expansion
//        Y[j] |= X[k]   for every edge (k,j) of A,    X[k] = bitmask of the F rows holding k,
// i.e. ONE pass over A serves up to 64*W frontier rows at once.  F lives as a vertex-major
// bit-matrix (DevBits) between hops; the sorted CSR the reference iterates
// (matrix.rs:1471-1605) is materialised only when something observes it.
//   push : active vertices only, load-balanced flat expansion, RED.OR.64 into Y
//   pull : stream A' (CSC of A) row by row, gather X[k], plain stores -- no atomics
// Algorithmic bytes per hop (pull): 4*nnz(A') + 8*(n+1) + 8*W*nnz(A') gathers (L2) + 8*W*n.
#include "common.cuh"
#include "ops.cuh"
#include <cub/block/block_reduce.cuh>
#include <algorithm>
#include <cstring>

namespace b200 {

// how the lanes of a warp share a W-word vertex record: a lane moves at most 32 bytes per instruction (LDG.256 / STG.256)
template <int W> struct PullCfg {
    static constexpr int WL = W >= 4 ? 4 : W;      // words per lane
    static constexpr int SPLIT = W / WL;           // lanes per vertex record
    static constexpr int VG = 8 / SPLIT;           // vertices gathered per 8-lane group and unroll step
};
static const u64 PUSH_CHUNK = 8192;
static const u64 LONG_ROW = 4096;   // pull: rows longer than this are split over several CTAs
static const u64 LONG_CHUNK = 8192;

static void require_natural(const DevBits &X, const char *what);

u32 bits_words_for(u64 nrows) {
    if (nrows == 0) return 1;
    u64 w = (nrows + 63) / 64;
    if (w > 16) return 0;
    u32 p = 1;
    while (p < w) p <<= 1;
    return p;
}

// ---------------------------------------------------------------------------- CSR -> bits
__global__ void k_bits_scatter(const u64 *__restrict__ Ap, const u32 *__restrict__ Aj, u64 nrows, u64 nnz,
                               u64 *__restrict__ X, u32 W) {
    u64 q = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    u64 stride = (u64)gridDim.x * blockDim.x;
    for (; q < nnz; q += stride) {
        // row = largest r with Ap[r] <= q
        u64 lo = 0, hi = nrows - 1;
        while (lo < hi) {
            u64 mid = (lo + hi + 1) >> 1;
            if (Ap[mid] <= q) lo = mid; else hi = mid - 1;
        }
        u32 col = Aj[q];
        atomicOr((unsigned long long *)&X[(u64)col * W + (lo >> 6)], 1ULL << (lo & 63));
    }
}

void bits_from_csr(const DevCSR &F, DevBits &X) {
    u32 W = bits_words_for(F.nrows);
    if (!W) throw GrbError(-8, "bit-frontier form needs <= 1024 rows");
    X.clear();
    X.nrows = F.nrows; X.ncols = F.ncols; X.W = W;
    X.w.alloc(F.ncols * W);
    X.w.zero();
    if (F.nnz)
        LAUNCH(k_bits_scatter, grid_for(F.nnz, 256, 1 << 16), 256, 0, F.p.ptr, F.j.ptr, F.nrows, F.nnz, X.w.ptr, W);
}

// ---------------------------------------------------------------------------- bits -> CSR
// Materialise the sorted CSR the reference's row iterator walks (matrix.rs:1471-1605) from the vertex-major
// bit-matrix.  A warp owns 32 consecutive vertices; a 5-stage shuffle butterfly transposes their 32x32 bit
// block so lane b holds "which of my 32 vertices are in frontier row b" (and row 32+b for the high half).
// Count pass: popcounts per (tile, row), laid out row-major-by-tile so ONE global exclusive scan yields final
// CSR positions.  Fill pass: each lane expands its rows' masks into a shared-memory list (u16 tile-local ids),
// then every row segment is copied out with coalesced stores -- ascending vertex id within each row.
static const u32 TILE_V = 1024;          // vertices per CTA tile
static const u32 TILE_THREADS = 512;     // 16 warps x 2 groups of 32 vertices

__device__ __forceinline__ u32 transpose32(u32 x, u32 lane) {
    // bit c of lane l  ->  bit l of lane c
    u32 y;
    y = __shfl_xor_sync(0xffffffffu, x, 16);
    x = (lane & 16) ? ((x & 0xFFFF0000u) | ((y & 0xFFFF0000u) >> 16)) : ((x & 0x0000FFFFu) | ((y & 0x0000FFFFu) << 16));
    y = __shfl_xor_sync(0xffffffffu, x, 8);
    x = (lane & 8) ? ((x & 0xFF00FF00u) | ((y & 0xFF00FF00u) >> 8)) : ((x & 0x00FF00FFu) | ((y & 0x00FF00FFu) << 8));
    y = __shfl_xor_sync(0xffffffffu, x, 4);
    x = (lane & 4) ? ((x & 0xF0F0F0F0u) | ((y & 0xF0F0F0F0u) >> 4)) : ((x & 0x0F0F0F0Fu) | ((y & 0x0F0F0F0Fu) << 4));
    y = __shfl_xor_sync(0xffffffffu, x, 2);
    x = (lane & 2) ? ((x & 0xCCCCCCCCu) | ((y & 0xCCCCCCCCu) >> 2)) : ((x & 0x33333333u) | ((y & 0x33333333u) << 2));
    y = __shfl_xor_sync(0xffffffffu, x, 1);
    x = (lane & 1) ? ((x & 0xAAAAAAAAu) | ((y & 0xAAAAAAAAu) >> 1)) : ((x & 0x55555555u) | ((y & 0x55555555u) << 1));
    return x;
}

// tc[(w*64 + row) * ntiles + tile] = number of vertices of the tile in frontier row 64w+row
__global__ void __launch_bounds__(TILE_THREADS)
k_bits_count(const u64 *__restrict__ X, u64 n, u32 W, u64 ntiles, u32 *__restrict__ tc) {
    __shared__ u32 wc[TILE_THREADS / 32][64];
    const u32 tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    u64 tile = blockIdx.x;
    u64 v0 = tile * TILE_V + (u64)warp * 64;
    for (u32 w = 0; w < W; w++) {
        u32 clo = 0, chi = 0;
#pragma unroll
        for (u32 g = 0; g < 2; g++) {
            u64 v = v0 + g * 32 + lane;
            u64 word = (v < n) ? X[v * W + w] : 0ULL;
            if (__ballot_sync(0xffffffffu, word != 0ULL)) {
                clo += __popc(transpose32((u32)word, lane));
                chi += __popc(transpose32((u32)(word >> 32), lane));
            }
        }
        wc[warp][lane] = clo;
        wc[warp][lane + 32] = chi;
        __syncthreads();
        if (tid < 64) {
            u32 s = 0;
#pragma unroll
            for (u32 q = 0; q < TILE_THREADS / 32; q++) s += wc[q][tid];
            tc[((u64)w * 64 + tid) * ntiles + tile] = s;
        }
        __syncthreads();
    }
}

// Count pass with vertical counters (count_kernel = 1, default).  The counts per (frontier row, tile) are a positional
// popcount: for each of the 64 bit positions of a word column, how many of the tile's 1024 words have that bit set.  No bit
// transpose per 32 vertices is needed for that: a lane adds the words of 32 vertices (v = lane, lane + 32, ...) into a
// carry-save adder tree -- 31 CSAs of two LOP3 per 32-bit half leave six bit-planes (weights 1, 2, .., 32) -- and only the
// planes are transposed across the warp (12 butterflies per word column instead of 64): ~4x fewer instructions than
// transposing every word.  One warp per (tile, group of up to 4 word columns); loads are 8*WL contiguous bytes per lane.
__device__ __forceinline__ void csa(u64 &hi, u64 &lo, u64 a, u64 b, u64 c) {   // a + b + c = 2 * hi + lo, bitwise
    const u64 u = a ^ b;
    hi = (a & b) | (u & c);
    lo = u ^ c;
}
template <int WL> __device__ __forceinline__ void ld_part(u64 (&x)[WL], const u64 *p) {
    if constexpr (WL == 1) { x[0] = __ldg(p); }
    else if constexpr (WL == 2) { ulonglong2 v = __ldg(reinterpret_cast<const ulonglong2 *>(p)); x[0] = v.x; x[1] = v.y; }
    else { u64x4 v = ld_v4(p); x[0] = v.a; x[1] = v.b; x[2] = v.c; x[3] = v.d; }
}
template <int W>
__global__ void __launch_bounds__(128)
k_bits_count_csa(const u64 *__restrict__ X, u64 n, u64 ntiles, u32 *__restrict__ tc) {
    constexpr int WL = PullCfg<W>::WL, SPLIT = PullCfg<W>::SPLIT;
    const u32 lane = threadIdx.x & 31;
    const u64 wid = ((u64)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const u64 tile = wid / SPLIT;
    const u32 h = (u32)(wid % SPLIT);
    if (tile >= ntiles) return;
    const u64 vbase = tile * TILE_V;
    // plane[p][w]: bit b = bit p of (number of this lane's 32 words of column w that have bit b set)
    u64 plane[6][WL];
    u64 pend[5][WL];            // carry waiting for its partner at level p + 1 (weight 2^(p+1))
#pragma unroll
    for (int p = 0; p < 6; p++)
#pragma unroll
        for (int w = 0; w < WL; w++) plane[p][w] = 0;
#pragma unroll
    for (int i = 0; i < 16; i++) {          // 16 pairs of vertices: v = lane + 32 * (2i), lane + 32 * (2i + 1)
        u64 xa[WL], xb[WL];
        const u64 va = vbase + lane + 64 * (u64)i, vb_ = va + 32;
        if (va < n) ld_part<WL>(xa, X + va * W + h * WL); else { for (int w = 0; w < WL; w++) xa[w] = 0; }
        if (vb_ < n) ld_part<WL>(xb, X + vb_ * W + h * WL); else { for (int w = 0; w < WL; w++) xb[w] = 0; }
#pragma unroll
        for (int w = 0; w < WL; w++) {
            u64 c;
            csa(c, plane[0][w], plane[0][w], xa[w], xb[w]);      // carry of weight 2
            // binary counter of pending carries: level L combines two carries of weight 2^L with the running plane
#pragma unroll
            for (int L = 1; L < 6; L++) {
                if ((i >> (L - 1)) & 1) {                        // second carry of this weight: combine, propagate
                    if (L < 5) { u64 c2; csa(c2, plane[L][w], plane[L][w], pend[L - 1][w], c); c = c2; }
                    else plane[5][w] ^= 0;                       // (unreachable: i < 16)
                } else { pend[L - 1][w] = c; break; }
            }
            if (i == 15) plane[5][w] = c;                        // the single carry of weight 32
        }
    }
    // planes -> counts: after transpose32 of one half-plane lane b holds bit b of all 32 lanes
#pragma unroll
    for (int w = 0; w < WL; w++) {
        u32 clo = 0, chi = 0;
#pragma unroll
        for (int p = 0; p < 6; p++) {
            clo += __popc(transpose32((u32)plane[p][w], lane)) << p;
            chi += __popc(transpose32((u32)(plane[p][w] >> 32), lane)) << p;
        }
        const u64 col = (u64)(h * WL + w) * 64;
        tc[(col + lane) * ntiles + tile] = clo;
        tc[(col + 32 + lane) * ntiles + tile] = chi;
    }
}

// Row-per-warp materialise (fill_kernel = 1, default).  Phase A transposes the tile exactly like k_bits_count and parks
// the 64 x 32 row masks in shared memory; phase B gives each warp whole rows: lane l owns the 32 vertices of mask word l
// (all lanes see the same row, so their populations are alike and the expansion loop barely diverges), a warp scan places
// the lanes, the ids go to a warp-private list aligned with the destination modulo 4 entries, and the list leaves as
// 16-byte stores.  One block barrier per word column (the masks are double-buffered); no block-wide prefix.
static const u32 FILL2_TSTRIDE = 33;                       // padded row stride of the mask tile (bank-conflict free)
static const u32 FILL2_LIST = TILE_V + 4;                  // entries per warp list (+3 alignment slack, rounded)
static const size_t FILL2_SMEM = 2 * 64 * FILL2_TSTRIDE * sizeof(u32) + (TILE_THREADS / 32) * FILL2_LIST * sizeof(unsigned short);
__global__ void __launch_bounds__(TILE_THREADS)
k_bits_fill_rows(const u64 *__restrict__ X, u64 n, u32 W, u64 ntiles, const u64 *__restrict__ off, u32 *__restrict__ Cj) {
    extern __shared__ __align__(16) unsigned char fill2_smem[];
    u32 *T = reinterpret_cast<u32 *>(fill2_smem);                                   // [2][64][33]
    const u32 tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    unsigned short *L = reinterpret_cast<unsigned short *>(fill2_smem + 2 * 64 * FILL2_TSTRIDE * sizeof(u32)) + warp * FILL2_LIST;
    const u32 NW = TILE_THREADS / 32;
    const u64 strm = policy_stream();
    const u64 tile = blockIdx.x;
    const u64 vbase = tile * TILE_V;
    const u32 vb = (u32)vbase;
    for (u32 w = 0; w < W; w++) {
        u32 *Tw = T + (w & 1) * 64 * FILL2_TSTRIDE;
#pragma unroll
        for (u32 g = 0; g < 2; g++) {
            u64 v = vbase + (u64)warp * 64 + g * 32 + lane;
            u64 word = (v < n) ? X[v * W + w] : 0ULL;
            u32 tl = 0, th = 0;
            if (__ballot_sync(0xffffffffu, word != 0ULL)) {
                tl = transpose32((u32)word, lane);
                th = transpose32((u32)(word >> 32), lane);
            }
            Tw[lane * FILL2_TSTRIDE + 2 * warp + g] = tl;
            Tw[(lane + 32) * FILL2_TSTRIDE + 2 * warp + g] = th;
        }
        __syncthreads();
        for (u32 r = warp; r < 64; r += NW) {
            u32 m = Tw[r * FILL2_TSTRIDE + lane];
            u32 c = __popc(m), incl = c;
#pragma unroll
            for (u32 d = 1; d < 32; d <<= 1) { u32 t = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += t; }
            const u32 cnt = __shfl_sync(0xffffffffu, incl, 31);
            if (cnt == 0) continue;
            const u64 g0 = off[((u64)w * 64 + r) * ntiles + tile];
            const u32 a = (u32)(g0 & 3);
            u32 o = a + incl - c;
            const u32 idb = lane * 32;
            while (m) { u32 bit = __ffs(m) - 1; L[o++] = (unsigned short)(idb + bit); m &= m - 1; }
            __syncwarp();
            u32 *dst = Cj + (g0 - a);                       // 16-byte aligned; dst[i] <-> L[i] for i in [a, a + cnt)
            const u32 total = a + cnt;
            const u32 head_end = a ? (total < 4 ? total : 4) : 0;
            if (lane >= a && lane < head_end) st_u32_stream(dst + lane, vb + L[lane], strm);
            const u32 nfull = total >> 2;
            const uint2 *L2 = reinterpret_cast<const uint2 *>(L);
            uint4 *dv = reinterpret_cast<uint4 *>(dst);
            for (u32 i = (a ? 1 : 0) + lane; i < nfull; i += 32) {
                uint2 p = L2[i];
                uint4 q;
                q.x = vb + (p.x & 0xFFFFu); q.y = vb + (p.x >> 16); q.z = vb + (p.y & 0xFFFFu); q.w = vb + (p.y >> 16);
                __stcs(dv + i, q);
            }
            const u32 tail = nfull * 4 > head_end ? nfull * 4 : head_end;
            if (tail + lane < total) st_u32_stream(dst + tail + lane, vb + L[tail + lane], strm);
            __syncwarp();
        }
    }
}

// Row-per-warp materialise over 2048-vertex tiles (fill_kernel = 3, default).  Same two phases as k_bits_fill_rows; what changed is
// the bookkeeping per emitted entry, which is what bounds this kernel (issue-active 81 %, profiles/r1f): a (row, tile) pass now
// covers 2048 vertices (each lane owns a 64-bit mask), a warp scans the counts of its rows in PAIRS (two 16-bit fields per
// shuffle), the four output offsets of a word column are fetched before the scans so their latency overlaps, and the staged ids
// keep only their low 16 bits -- the high half is merged with one PRMT on the way out.  ~1.0 instructions per entry against 1.6.
static const u32 F3_TILE = 2048, F3_THREADS = 512, F3_STRIDE = 66, F3_LIST = F3_TILE + 8;
static const size_t F3_SMEM = 2 * 64 * F3_STRIDE * sizeof(u32) + (F3_THREADS / 32) * F3_LIST * sizeof(unsigned short);
__global__ void __launch_bounds__(F3_THREADS, 2)
k_bits_fill_v3(const u64 *__restrict__ X, u64 n, u32 W, u64 ntiles1k, const u64 *__restrict__ off, u32 *__restrict__ Cj) {
    extern __shared__ __align__(16) unsigned char f3_smem[];
    u32 *T = reinterpret_cast<u32 *>(f3_smem);                                      // [2][64][F3_STRIDE]
    const u32 tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    unsigned short *L = reinterpret_cast<unsigned short *>(f3_smem + 2 * 64 * F3_STRIDE * sizeof(u32)) + warp * F3_LIST;
    constexpr u32 NW = F3_THREADS / 32, RPW = 64 / NW;                              // 16 warps, 4 rows per warp and word column
    const u64 strm = policy_stream();
    const u64 tile = blockIdx.x;
    const u64 vbase = tile * F3_TILE;
    const u32 lo16base = (u32)(vbase & 0xFFFFu), vbhi = (u32)(vbase >> 16);
    for (u32 w = 0; w < W; w++) {
        u32 *Tw = T + (w & 1) * 64 * F3_STRIDE;
        // offsets of this warp's rows: issued first, consumed after the scans
        u64 g0[RPW];
#pragma unroll
        for (u32 i = 0; i < RPW; i++) g0[i] = off[((u64)w * 64 + warp + NW * i) * ntiles1k + 2 * tile];
#pragma unroll
        for (u32 g = 0; g < 4; g++) {
            const u64 v = vbase + (u64)warp * 128 + g * 32 + lane;
            const u64 word = (v < n) ? X[v * W + w] : 0ULL;
            u32 tl = 0, th = 0;
            if (__ballot_sync(0xffffffffu, word != 0ULL)) {
                tl = transpose32((u32)word, lane);
                th = transpose32((u32)(word >> 32), lane);
            }
            Tw[lane * F3_STRIDE + 4 * warp + g] = tl;
            Tw[(lane + 32) * F3_STRIDE + 4 * warp + g] = th;
        }
        __syncthreads();
        u64 m[RPW];
        u32 c[RPW];
#pragma unroll
        for (u32 i = 0; i < RPW; i++) {
            m[i] = *reinterpret_cast<const u64 *>(Tw + (warp + NW * i) * F3_STRIDE + 2 * lane);
            c[i] = __popcll(m[i]);
        }
        u32 pk[RPW / 2];
#pragma unroll
        for (u32 h = 0; h < RPW / 2; h++) pk[h] = c[2 * h] | (c[2 * h + 1] << 16);     // prefixes stay below 2^16 (tile = 2048)
#pragma unroll
        for (u32 d = 1; d < 32; d <<= 1) {
#pragma unroll
            for (u32 h = 0; h < RPW / 2; h++) { const u32 t = __shfl_up_sync(0xffffffffu, pk[h], d); if (lane >= d) pk[h] += t; }
        }
#pragma unroll
        for (u32 i = 0; i < RPW; i++) {
            const u32 field = (i & 1) ? (pk[i / 2] >> 16) : (pk[i / 2] & 0xFFFFu);
            const u32 cnt = __shfl_sync(0xffffffffu, field, 31);
            if (cnt == 0) continue;
            const u32 a = (u32)(g0[i] & 3);
            u32 o = a + field - c[i];
            u32 idb = lo16base + lane * 64;
            u32 lo = (u32)m[i], hi = (u32)(m[i] >> 32);
            while (lo) { const u32 bit = __ffs(lo) - 1; L[o++] = (unsigned short)(idb + bit); lo &= lo - 1; }
            idb += 32;
            while (hi) { const u32 bit = __ffs(hi) - 1; L[o++] = (unsigned short)(idb + bit); hi &= hi - 1; }
            __syncwarp();
            u32 *dst = Cj + (g0[i] - a);                    // 16-byte aligned; dst[k] <-> L[k] for k in [a, a + cnt)
            const u32 total = a + cnt;
            const u32 head_end = a ? (total < 4 ? total : 4) : 0;
            if (lane >= a && lane < head_end) st_u32_stream(dst + lane, (vbhi << 16) | L[lane], strm);
            const u32 nfull = total >> 2;
            const uint2 *L2 = reinterpret_cast<const uint2 *>(L);
            uint4 *dv = reinterpret_cast<uint4 *>(dst);
            for (u32 k = (a ? 1 : 0) + lane; k < nfull; k += 32) {
                const uint2 p = L2[k];
                uint4 q;
                q.x = __byte_perm(p.x, vbhi, 0x5410); q.y = __byte_perm(p.x, vbhi, 0x5432);
                q.z = __byte_perm(p.y, vbhi, 0x5410); q.w = __byte_perm(p.y, vbhi, 0x5432);
                __stcs(dv + k, q);
            }
            const u32 tail = nfull * 4 > head_end ? nfull * 4 : head_end;
            if (tail + lane < total) st_u32_stream(dst + tail + lane, (vbhi << 16) | L[tail + lane], strm);
            __syncwarp();
        }
    }
}

// Materialise with the transposed masks kept (fill_kernel = 2, default).  k_bits_count_keep counts like k_bits_count and also
// writes the 64 x 32 row masks of every (word column, tile) to global memory, coalesced (8 KB per tile); k_bits_fill_masks is
// then phase B of k_bits_fill_rows alone: a warp reads one row's 32 mask words with one 128-byte load -- no second transpose,
// no shared tile, no block barrier.  The masks cost one extra write + read of the bit-matrix (2 x 8*W*n bytes), cheap for
// kernels that are instruction-bound, not DRAM-bound.
__global__ void __launch_bounds__(TILE_THREADS)
k_bits_count_keep(const u64 *__restrict__ X, u64 n, u32 W, u64 ntiles, u32 *__restrict__ tc, u32 *__restrict__ Tg) {
    __shared__ __align__(16) u32 T[64 * FILL2_TSTRIDE];
    const u32 tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const u64 tile = blockIdx.x;
    const u64 vbase = tile * TILE_V;
    for (u32 w = 0; w < W; w++) {
#pragma unroll
        for (u32 g = 0; g < 2; g++) {
            u64 v = vbase + (u64)warp * 64 + g * 32 + lane;
            u64 word = (v < n) ? X[v * W + w] : 0ULL;
            u32 tl = 0, th = 0;
            if (__ballot_sync(0xffffffffu, word != 0ULL)) {
                tl = transpose32((u32)word, lane);
                th = transpose32((u32)(word >> 32), lane);
            }
            T[lane * FILL2_TSTRIDE + 2 * warp + g] = tl;
            T[(lane + 32) * FILL2_TSTRIDE + 2 * warp + g] = th;
        }
        __syncthreads();
        {   // thread (r, c4): 4 consecutive mask words of row r -> one 16-byte store; 8 threads cover the row
            const u32 r = tid >> 3, c4 = (tid & 7) * 4;
            uint4 q;
            q.x = T[r * FILL2_TSTRIDE + c4]; q.y = T[r * FILL2_TSTRIDE + c4 + 1];
            q.z = T[r * FILL2_TSTRIDE + c4 + 2]; q.w = T[r * FILL2_TSTRIDE + c4 + 3];
            u32 c = __popc(q.x) + __popc(q.y) + __popc(q.z) + __popc(q.w);
            c += __shfl_xor_sync(0xffffffffu, c, 1);
            c += __shfl_xor_sync(0xffffffffu, c, 2);
            c += __shfl_xor_sync(0xffffffffu, c, 4);
            if ((tid & 7) == 0) tc[((u64)w * 64 + r) * ntiles + tile] = c;
            *reinterpret_cast<uint4 *>(Tg + (((u64)w * ntiles + tile) * 64 + r) * 32 + c4) = q;
        }
        __syncthreads();
    }
}

static const u32 FILLM_WARPS = 8;                          // warps per CTA of k_bits_fill_masks (one list each)
__global__ void __launch_bounds__(FILLM_WARPS * 32)
k_bits_fill_masks(const u32 *__restrict__ Tg, u32 W, u64 ntiles, const u64 *__restrict__ off, u32 *__restrict__ Cj) {
    __shared__ __align__(16) unsigned short lists[FILLM_WARPS][FILL2_LIST];
    const u32 warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    unsigned short *L = lists[warp];
    const u64 strm = policy_stream();
    const u64 ntasks = (u64)W * ntiles * 64;               // task = (word column, tile, row), masks stored in that order
    const u64 nwarps = (u64)gridDim.x * FILLM_WARPS;
    u64 task = (u64)blockIdx.x * FILLM_WARPS + warp;
    u32 mnext = task < ntasks ? __ldg(Tg + task * 32 + lane) : 0;
    for (; task < ntasks; task += nwarps) {
        u32 m = mnext;
        if (task + nwarps < ntasks) mnext = __ldg(Tg + (task + nwarps) * 32 + lane);   // next task's masks in flight
        u32 c = __popc(m), incl = c;
#pragma unroll
        for (u32 d = 1; d < 32; d <<= 1) { u32 t = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += t; }
        const u32 cnt = __shfl_sync(0xffffffffu, incl, 31);
        if (cnt == 0) continue;
        const u32 r = (u32)(task & 63);
        const u64 wt = task >> 6, w = wt / ntiles, tile = wt - w * ntiles;
        const u32 vb = (u32)(tile * TILE_V);
        const u64 g0 = off[(w * 64 + r) * ntiles + tile];
        const u32 a = (u32)(g0 & 3);
        u32 o = a + incl - c;
        const u32 idb = lane * 32;
        while (m) { u32 bit = __ffs(m) - 1; L[o++] = (unsigned short)(idb + bit); m &= m - 1; }
        __syncwarp();
        u32 *dst = Cj + (g0 - a);                           // 16-byte aligned; dst[i] <-> L[i] for i in [a, a + cnt)
        const u32 total = a + cnt;
        const u32 head_end = a ? (total < 4 ? total : 4) : 0;
        if (lane >= a && lane < head_end) st_u32_stream(dst + lane, vb + L[lane], strm);
        const u32 nfull = total >> 2;
        const uint2 *L2 = reinterpret_cast<const uint2 *>(L);
        uint4 *dv = reinterpret_cast<uint4 *>(dst);
        for (u32 i = (a ? 1 : 0) + lane; i < nfull; i += 32) {
            uint2 p = L2[i];
            uint4 q;
            q.x = vb + (p.x & 0xFFFFu); q.y = vb + (p.x >> 16); q.z = vb + (p.y & 0xFFFFu); q.w = vb + (p.y >> 16);
            __stcs(dv + i, q);
        }
        const u32 tail = nfull * 4 > head_end ? nfull * 4 : head_end;
        if (tail + lane < total) st_u32_stream(dst + tail + lane, vb + L[tail + lane], strm);
        __syncwarp();
    }
}

__global__ void __launch_bounds__(TILE_THREADS)
k_bits_fill(const u64 *__restrict__ X, u64 n, u32 W, u64 ntiles, const u64 *__restrict__ off, u32 *__restrict__ Cj,
            u32 cap) {
    extern __shared__ unsigned short list[];             // `cap` tile-local ids (sized from the average tile density)
    __shared__ unsigned short wp[TILE_THREADS / 32][64];  // per-warp exclusive prefix, per row
    __shared__ u32 sbase[65];                            // row segment starts inside `list`
    __shared__ u64 goff[64];                             // row segment starts in the output
    const u32 tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const u32 NW = TILE_THREADS / 32;
    const u64 strm = policy_stream();
    u64 tile = blockIdx.x;
    u64 vbase = tile * TILE_V;
    for (u32 w = 0; w < W; w++) {
        u32 tlo[2], thi[2];
        u32 clo = 0, chi = 0;
#pragma unroll
        for (u32 g = 0; g < 2; g++) {
            u64 v = vbase + (u64)warp * 64 + g * 32 + lane;
            u64 word = (v < n) ? X[v * W + w] : 0ULL;
            tlo[g] = 0; thi[g] = 0;
            if (__ballot_sync(0xffffffffu, word != 0ULL)) {
                tlo[g] = transpose32((u32)word, lane);
                thi[g] = transpose32((u32)(word >> 32), lane);
            }
            clo += __popc(tlo[g]);
            chi += __popc(thi[g]);
        }
        wp[warp][lane] = (unsigned short)clo;
        wp[warp][lane + 32] = (unsigned short)chi;
        __syncthreads();
        if (tid < 64) {
            u32 run = 0;
#pragma unroll
            for (u32 q = 0; q < NW; q++) { u32 t = wp[q][tid]; wp[q][tid] = (unsigned short)run; run += t; }
            sbase[tid + 1] = run; // totals for now
            goff[tid] = off[((u64)w * 64 + tid) * ntiles + tile];
        }
        __syncthreads();
        if (tid == 0) {
            u32 run = 0;
            sbase[0] = 0;
            for (u32 r = 0; r < 64; r++) { u32 t = sbase[r + 1]; sbase[r + 1] = run + t; run += t; }
        }
        __syncthreads();
        const bool staged = sbase[64] <= cap;   // denser-than-provisioned tiles write straight to global memory
        if (!staged) {
            u64 o = goff[lane] + wp[warp][lane];
#pragma unroll
            for (u32 g = 0; g < 2; g++) {
                u32 m = tlo[g];
                u64 idb = vbase + warp * 64 + g * 32;
                while (m) { u32 bit = __ffs(m) - 1; Cj[o++] = (u32)(idb + bit); m &= m - 1; }
            }
            o = goff[lane + 32] + wp[warp][lane + 32];
#pragma unroll
            for (u32 g = 0; g < 2; g++) {
                u32 m = thi[g];
                u64 idb = vbase + warp * 64 + g * 32;
                while (m) { u32 bit = __ffs(m) - 1; Cj[o++] = (u32)(idb + bit); m &= m - 1; }
            }
            __syncthreads();
            continue;
        }
        // expand: lane b owns rows b (low half) and 32+b (high half) of this warp's 64 vertices
        {
            u32 o = sbase[lane] + wp[warp][lane];
#pragma unroll
            for (u32 g = 0; g < 2; g++) {
                u32 m = tlo[g];
                u32 idb = warp * 64 + g * 32;
                while (m) { u32 bit = __ffs(m) - 1; list[o++] = (unsigned short)(idb + bit); m &= m - 1; }
            }
            o = sbase[lane + 32] + wp[warp][lane + 32];
#pragma unroll
            for (u32 g = 0; g < 2; g++) {
                u32 m = thi[g];
                u32 idb = warp * 64 + g * 32;
                while (m) { u32 bit = __ffs(m) - 1; list[o++] = (unsigned short)(idb + bit); m &= m - 1; }
            }
        }
        __syncthreads();
        // coalesced copy-out: warp q handles rows q, q+NW, ...
        for (u32 r = warp; r < 64; r += NW) {
            u32 s = sbase[r], cnt = sbase[r + 1] - s;
            u32 *dst = Cj + goff[r];
            for (u32 i = lane; i < cnt; i += 32) st_u32_stream(dst + i, (u32)(vbase + list[s + i]), strm);
        }
        __syncthreads();
    }
}

__global__ void k_bits_rowptr(const u64 *__restrict__ off, u64 nrows, u64 ntiles, u64 *__restrict__ Cp) {
    u64 r = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (r <= nrows) Cp[r] = off[r * ntiles];
}

void bits_to_csr(const DevBits &X, DevCSR &C) {
    require_natural(X, "bits_to_csr");
    u64 n = X.ncols;
    u32 W = X.W;
    C.clear();
    C.nrows = X.nrows; C.ncols = n;
    C.p.alloc(X.nrows + 1);
    if (n == 0) { C.p.zero(); C.nnz = 0; return; }
    u64 ntiles = (n + TILE_V - 1) / TILE_V;
    u64 ncnt = (u64)64 * W * ntiles;
    DevBuf<u32> tc(ncnt + 1);
    DevBuf<u64> off(ncnt + 1);
    const bool keep_masks = ctx().opt_fill_kernel == 2 && ctx().opt_fill_cap <= 0;
    DevBuf<u32> masks;
    if (keep_masks) masks.alloc((u64)W * ntiles * 64 * 32);
    {
        TimedScope ts(TK_BITS_COUNT, 8ULL * W * n + (keep_masks ? 8ULL * W * ntiles * TILE_V : 0));
        if (keep_masks) LAUNCH(k_bits_count_keep, (u32)ntiles, TILE_THREADS, 0, X.w.ptr, n, W, ntiles, tc.ptr, masks.ptr);
        else if (ctx().opt_count_kernel == 1) {
            const u64 nwarps = ntiles * (W >= 4 ? W / 4 : 1);
            const u32 g = (u32)((nwarps + 3) / 4);
            switch (W) {
            case 1: LAUNCH((k_bits_count_csa<1>), g, 128, 0, X.w.ptr, n, ntiles, tc.ptr); break;
            case 2: LAUNCH((k_bits_count_csa<2>), g, 128, 0, X.w.ptr, n, ntiles, tc.ptr); break;
            case 4: LAUNCH((k_bits_count_csa<4>), g, 128, 0, X.w.ptr, n, ntiles, tc.ptr); break;
            case 8: LAUNCH((k_bits_count_csa<8>), g, 128, 0, X.w.ptr, n, ntiles, tc.ptr); break;
            default: LAUNCH((k_bits_count_csa<16>), g, 128, 0, X.w.ptr, n, ntiles, tc.ptr); break;
            }
        }
        else LAUNCH(k_bits_count, (u32)ntiles, TILE_THREADS, 0, X.w.ptr, n, W, ntiles, tc.ptr);
    }
    CUDA_TRY(cudaMemsetAsync(tc.ptr + ncnt, 0, sizeof(u32), stream()));
    exclusive_scan_u32_to_u64(tc.ptr, off.ptr, ncnt + 1);
    LAUNCH(k_bits_rowptr, grid_for(X.nrows + 1, 256), 256, 0, off.ptr, X.nrows, ntiles, C.p.ptr);
    u64 nnz = read_scalar(off.ptr + X.nrows * ntiles);
    C.nnz = nnz;
    C.j.alloc(nnz);
    if (nnz) {
        static bool attr_set = false;
        if (!attr_set) {
            CUDA_TRY(cudaFuncSetAttribute(k_bits_fill, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(64 * TILE_V * sizeof(unsigned short))));
            CUDA_TRY(cudaFuncSetAttribute(k_bits_fill_rows, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FILL2_SMEM));
            CUDA_TRY(cudaFuncSetAttribute(k_bits_fill_v3, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)F3_SMEM));
            attr_set = true;
        }
        TimedScope ts(TK_BITS_FILL, 8ULL * W * n + 4 * nnz);
        if (keep_masks) {
            LAUNCH(k_bits_fill_masks, (u32)ctx().num_sms * 8, FILLM_WARPS * 32, 0, masks.ptr, W, ntiles, off.ptr, C.j.ptr);
        } else if (ctx().opt_fill_kernel == 3 && ctx().opt_fill_cap <= 0 && W <= 8) {
            // (at W = 16 the 2048-vertex tile's 128 KB of frontier words per CTA no longer survives in L1 next to 100 KB of shared
            // memory and the word columns are re-fetched from L2: 8.7 vs 6.2 ms -- the 1024-vertex kernel below takes W = 16)
            LAUNCH(k_bits_fill_v3, (u32)((n + F3_TILE - 1) / F3_TILE), F3_THREADS, F3_SMEM, X.w.ptr, n, W, ntiles, off.ptr, C.j.ptr);
        } else if ((ctx().opt_fill_kernel == 1 || ctx().opt_fill_kernel == 3) && ctx().opt_fill_cap <= 0) {
            LAUNCH(k_bits_fill_rows, (u32)ntiles, TILE_THREADS, FILL2_SMEM, X.w.ptr, n, W, ntiles, off.ptr, C.j.ptr);
        } else {
            // staging capacity from the average (tile, word) population, 1.5x headroom, 8K-entry steps: sparse frontiers
            // get small shared-memory footprints and therefore several resident CTAs per SM
            u64 avg = nnz / (ntiles * W) + 1;
            u32 cap = 8192;
            while (cap < 65536 && (u64)cap < avg + avg / 2) cap += 8192;
            if (ctx().opt_fill_cap > 0) cap = (u32)ctx().opt_fill_cap; // test hook: force the direct-write path
            const size_t smem = (size_t)cap * sizeof(unsigned short);
            LAUNCH(k_bits_fill, (u32)ntiles, TILE_THREADS, smem, X.w.ptr, n, W, ntiles, off.ptr, C.j.ptr, cap);
        }
    }
}

// ---------------------------------------------------------------------------- diagonal operand = column filter
// Label matrices are n x n diagonal (graph.rs:1191): F * L keeps column j of F iff L(j,j) is present -- an elementwise pass
// over the bit-matrix instead of a hop (label-filter fusion).
__global__ void __launch_bounds__(256) k_is_diagonal(const u64 *__restrict__ p, const u32 *__restrict__ j, u64 n, u32 *__restrict__ bad) {
    u64 r = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    u64 stride = (u64)gridDim.x * blockDim.x;
    bool b = false;
    for (; r < n; r += stride) {
        u64 s = p[r], e = p[r + 1];
        if (e - s > 1 || (e - s == 1 && j[s] != (u32)r)) b = true;
    }
    if (__any_sync(0xffffffffu, b) && (threadIdx.x & 31) == 0) atomicOr(bad, 1u);
}
bool csr_is_diagonal(const DevCSR &A) {
    if (A.nrows != A.ncols || A.nnz > A.nrows) return false;
    if (A.nrows == 0) return true;
    DevBuf<u32> bad(1);
    bad.zero();
    LAUNCH(k_is_diagonal, grid_for(A.nrows, 256, 148 * 16), 256, 0, A.p.ptr, A.j.ptr, A.nrows, bad.ptr);
    return read_scalar(bad.ptr) == 0;
}
__global__ void __launch_bounds__(256)
k_bits_diag(const u64 *__restrict__ X, const u64 *__restrict__ Ap, u64 n, u32 W, u64 *__restrict__ Y, u64 *__restrict__ flops) {
    typedef cub::BlockReduce<u64, 256> Red;
    __shared__ typename Red::TempStorage ts;
    u64 t = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    const u64 stride = (u64)gridDim.x * blockDim.x, total = n * W;
    u64 c = 0;
    for (; t < total; t += stride) {
        const u64 v = t / W;
        const u64 x = (Ap[v + 1] > Ap[v]) ? X[t] : 0ULL;
        Y[t] = x;
        c += __popcll(x);
    }
    u64 tot = Red(ts).Sum(c);
    if (threadIdx.x == 0 && tot) atomicAdd((unsigned long long *)flops, tot);
}
void bits_diag(const DevBits &X, const DevCSR &A, DevBits &Y, u64 *flops_out) {
    require_natural(X, "bits_diag");
    if (X.ncols != A.nrows) throw GrbError(-6, "mxm: inner dimensions differ");
    Y.clear();
    Y.nrows = X.nrows; Y.ncols = A.ncols; Y.W = X.W;
    Y.w.alloc(A.ncols * X.W);
    DevBuf<u64> fl(1);
    fl.zero();
    if (A.nrows) LAUNCH(k_bits_diag, grid_for(A.nrows * X.W, 256, 148 * 16), 256, 0, X.w.ptr, A.p.ptr, A.nrows, X.W, Y.w.ptr, fl.ptr);
    if (flops_out) *flops_out = read_scalar(fl.ptr);
}

// ---------------------------------------------------------------------------- row-major bitmap (GxB_BITMAP-like export)
// out[row * wpr + (v >> 6)] bit (v & 63) <=> X(row, v).  Same 64 x 1024 transposed tile as the materialise kernels, written
// out as 128-byte row segments; rows >= nrows (padding of the last word column) are dropped.
__global__ void __launch_bounds__(TILE_THREADS)
k_bits_rowmajor(const u64 *__restrict__ X, u64 n, u32 W, u64 nrows, u64 wpr, u64 *__restrict__ out) {
    __shared__ u32 T[2][64 * FILL2_TSTRIDE];
    const u32 tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const u64 tile = blockIdx.x;
    const u64 vbase = tile * TILE_V;
    const u64 strm = policy_stream();
    for (u32 w = 0; w < W; w++) {
        if ((u64)w * 64 >= nrows) break;
        u32 *Tw = T[w & 1];
#pragma unroll
        for (u32 g = 0; g < 2; g++) {
            u64 v = vbase + (u64)warp * 64 + g * 32 + lane;
            u64 word = (v < n) ? X[v * W + w] : 0ULL;
            u32 tl = 0, th = 0;
            if (__ballot_sync(0xffffffffu, word != 0ULL)) {
                tl = transpose32((u32)word, lane);
                th = transpose32((u32)(word >> 32), lane);
            }
            Tw[lane * FILL2_TSTRIDE + 2 * warp + g] = tl;
            Tw[(lane + 32) * FILL2_TSTRIDE + 2 * warp + g] = th;
        }
        __syncthreads();
#pragma unroll
        for (u32 h = 0; h < 2; h++) {
            u32 r = (tid >> 4) + 32 * h, c = tid & 15;
            u64 row = (u64)w * 64 + r, col = tile * (TILE_V / 64) + c;
            if (row < nrows && col < wpr) {
                u64 val = (u64)Tw[r * FILL2_TSTRIDE + 2 * c] | ((u64)Tw[r * FILL2_TSTRIDE + 2 * c + 1] << 32);
                st_u64_stream(out + row * wpr + col, val, strm);
            }
        }
    }
}

void bits_to_rowmajor(const DevBits &X, u64 *out, u64 wpr) {
    require_natural(X, "bits_to_rowmajor");
    u64 n = X.ncols;
    if (!n || !X.nrows) return;
    u64 ntiles = (n + TILE_V - 1) / TILE_V;
    LAUNCH(k_bits_rowmajor, (u32)ntiles, TILE_THREADS, 0, X.w.ptr, n, X.W, X.nrows, wpr, out);
}

// CSR -> row-major bitmap for matrices that are not in frontier form (out must be zeroed); one warp per row
__global__ void k_csr_to_bitmap(const u64 *__restrict__ p, const u32 *__restrict__ j, u64 nrows, u64 wpr, u64 *__restrict__ out) {
    u64 row = ((u64)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    u32 lane = threadIdx.x & 31;
    if (row >= nrows) return;
    for (u64 q = p[row] + lane; q < p[row + 1]; q += 32) {
        u32 c = j[q];
        atomicOr((unsigned long long *)(out + row * wpr + (c >> 6)), 1ULL << (c & 63));
    }
}
void csr_to_rowmajor(const DevCSR &A, u64 *out, u64 wpr) {
    if (!A.nrows || !A.nnz) return;
    LAUNCH(k_csr_to_bitmap, grid_for(A.nrows * 32, 256), 256, 0, A.p.ptr, A.j.ptr, A.nrows, wpr, out);
}

// ---------------------------------------------------------------------------- reductions
__global__ void __launch_bounds__(256) k_popc_sum(const u64 *__restrict__ w, u64 n, u64 *__restrict__ out) {
    typedef cub::BlockReduce<u64, 256> Red;
    __shared__ typename Red::TempStorage ts;
    u64 t = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    u64 stride = (u64)gridDim.x * blockDim.x;
    u64 s = 0;
    for (; t < n; t += stride) s += __popcll(w[t]);
    u64 tot = Red(ts).Sum(s);
    if (threadIdx.x == 0 && tot) atomicAdd((unsigned long long *)out, tot);
}

u64 bits_nvals(const DevBits &X) {
    DevBuf<u64> out(1);
    out.zero();
    u64 n = X.ncols * X.W;
    if (n) LAUNCH(k_popc_sum, grid_for(n, 256, 148 * 8), 256, 0, X.w.ptr, n, out.ptr);
    return read_scalar(out.ptr);
}

// flops (mxm sense) = sum_k popc(X[k]) * deg_A(k); edges = sum_{k active} deg_A(k); flag[k] = active
__global__ void __launch_bounds__(256)
k_bits_flops(const u64 *__restrict__ X, u32 W, u64 n, const u64 *__restrict__ Ap, u32 *__restrict__ flag,
             u64 *__restrict__ out /* [0]=flops [1]=edges */) {
    typedef cub::BlockReduce<u64, 256> Red;
    __shared__ typename Red::TempStorage ts, ts2;
    u64 k = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    u64 stride = (u64)gridDim.x * blockDim.x;
    u64 fl = 0, ed = 0;
    for (; k < n; k += stride) {
        u32 pc = 0;
        for (u32 w = 0; w < W; w++) pc += __popcll(X[k * W + w]);
        u64 d = Ap[k + 1] - Ap[k];
        flag[k] = (pc != 0 && d != 0) ? 1u : 0u;
        if (pc) { fl += (u64)pc * d; ed += d; }
    }
    u64 tf = Red(ts).Sum(fl);
    u64 te = Red(ts2).Sum(ed);
    if (threadIdx.x == 0) {
        if (tf) atomicAdd((unsigned long long *)&out[0], tf);
        if (te) atomicAdd((unsigned long long *)&out[1], te);
    }
}

// ---------------------------------------------------------------------------- push
__global__ void k_compact_active(const u32 *__restrict__ flag, const u64 *__restrict__ pos, u64 n,
                                 const u64 *__restrict__ Ap, u32 *__restrict__ act, u64 *__restrict__ deg,
                                 u64 *__restrict__ astart) {
    u64 k = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    u64 stride = (u64)gridDim.x * blockDim.x;
    for (; k < n; k += stride) {
        if (flag[k]) {
            u64 d = pos[k];
            act[d] = (u32)k;
            u64 s = Ap[k];
            deg[d] = Ap[k + 1] - s;
            astart[d] = s;
        }
    }
}

__device__ __forceinline__ u64 find_le64(const u64 *__restrict__ a, u64 lo, u64 hi, u64 target) {
    while (lo < hi) {
        u64 mid = (lo + hi + 1) >> 1;
        if (a[mid] <= target) lo = mid; else hi = mid - 1;
    }
    return lo;
}

template <int W>
__global__ void __launch_bounds__(256)
k_bits_push(const u32 *__restrict__ act, const u64 *__restrict__ cum, const u64 *__restrict__ astart, u64 nact,
            u64 total, const u32 *__restrict__ Aj, const u64 *__restrict__ X, u64 *__restrict__ Y) {
    __shared__ u64 s_e0, s_e1;
    const u32 tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    u64 lo = (u64)blockIdx.x * PUSH_CHUNK, hi = lo + PUSH_CHUNK;
    if (hi > total) hi = total;
    if (tid == 0) {
        s_e0 = find_le64(cum, 0, nact - 1, lo);
        s_e1 = find_le64(cum, 0, nact - 1, hi - 1);
    }
    __syncthreads();
    u64 e0 = s_e0, e1 = s_e1;
    for (u64 t0 = lo + (u64)warp * 32; t0 < hi; t0 += 8 * 32) {
        u64 e = find_le64(cum, e0, e1, t0);
        u64 t = t0 + lane;
        if (t < hi) {
            while (e < e1 && cum[e + 1] <= t) e++;
            u32 k = act[e];
            u32 col = Aj[astart[e] + (t - cum[e])];
#pragma unroll
            for (int w = 0; w < W; w++) {
                u64 xw = X[(u64)k * W + w];
                if (xw) atomicOr((unsigned long long *)&Y[(u64)col * W + w], xw);
            }
        }
    }
}

// ---------------------------------------------------------------------------- push straight from a CSR frontier
// Early hops of a traversal batch have a tiny frontier (F in CSR form, a few thousand entries).  Going through the
// bit-matrix there means O(n*W) passes that dwarf the real work (zeroing X, scanning it for flops / active vertices,
// compacting n flags); this path expands F's entries directly: entry (i, k) ORs bit i into Y[j] for every j in A(k,:).
__global__ void k_csr_entry_info(const u64 *__restrict__ Fp, const u32 *__restrict__ Fj, u64 nrows, const u64 *__restrict__ Ap,
                                 u32 *__restrict__ erow, u64 *__restrict__ edeg) {
    u64 warp = ((u64)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    u32 lane = threadIdx.x & 31;
    if (warp >= nrows) return;
    for (u64 q = Fp[warp] + lane; q < Fp[warp + 1]; q += 32) {
        u32 k = Fj[q];
        erow[q] = (u32)warp;
        edeg[q] = Ap[k + 1] - Ap[k];
    }
}
template <int W>
__global__ void __launch_bounds__(256)
k_csr_push(const u32 *__restrict__ erow, const u32 *__restrict__ Fj, const u64 *__restrict__ cum, u64 nent, u64 total,
           const u64 *__restrict__ Ap, const u32 *__restrict__ Aj, u64 *__restrict__ Y, const u32 *__restrict__ perm) {
    __shared__ u64 s_e0, s_e1;
    const u32 tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    u64 lo = (u64)blockIdx.x * PUSH_CHUNK, hi = lo + PUSH_CHUNK;
    if (hi > total) hi = total;
    if (tid == 0) {
        s_e0 = find_le64(cum, 0, nent - 1, lo);
        s_e1 = find_le64(cum, 0, nent - 1, hi - 1);
    }
    __syncthreads();
    u64 e0 = s_e0, e1 = s_e1;
    for (u64 t0 = lo + (u64)warp * 32; t0 < hi; t0 += 8 * 32) {
        u64 e = find_le64(cum, e0, e1, t0);
        u64 t = t0 + lane;
        if (t < hi) {
            while (e < e1 && cum[e + 1] <= t) e++;
            const u32 i = erow[e];
            u32 col = Aj[Ap[Fj[e]] + (t - cum[e])];
            if (perm) col = __ldg(perm + col);          // hot-set order: hubs of the next gather become neighbours (64 MB table, L2)
            atomicOr((unsigned long long *)&Y[(u64)col * W + (i >> 6)], 1ULL << (i & 63));
        }
    }
}
template <int W>
static void csr_push_impl(const DevCSR &F, const DevCSR &A, DevBits &Y, u64 flops, const u32 *erow, const u64 *cum, const u32 *perm) {
    u64 nchunks = (flops + PUSH_CHUNK - 1) / PUSH_CHUNK;
    TimedScope ts(TK_BITS_PUSH, 4 * flops + 16 * F.nnz + 8 * flops);
    LAUNCH((k_csr_push<W>), (u32)nchunks, 256, 0, erow, F.j.ptr, cum, F.nnz, flops, A.p.ptr, A.j.ptr, Y.w.ptr, perm);
}
// Y = F * A from the CSR form of F.  Returns false (nothing done) when the expansion is large enough that the
// direction-optimising bit-matrix hop should take it: flops * 4 > nnz(A), the same switch bits_hop applies to edges.
bool bits_push_from_csr(const DevCSR &F, const DevCSR &A, DevBits &Y, u64 *flops_out, const LongRows *lr) {
    if (F.ncols != A.nrows) throw GrbError(-6, "mxm: inner dimensions differ");
    const u32 W = bits_words_for(F.nrows);
    u64 flops = 0;
    DevBuf<u32> erow(F.nnz);
    DevBuf<u64> cum(F.nnz + 1);
    if (F.nnz) {
        LAUNCH(k_csr_entry_info, grid_for(F.nrows * 32, 256), 256, 0, F.p.ptr, F.j.ptr, F.nrows, A.p.ptr, erow.ptr, cum.ptr);
        CUDA_TRY(cudaMemsetAsync(cum.ptr + F.nnz, 0, sizeof(u64), stream()));
        exclusive_scan_u64(cum.ptr, cum.ptr, F.nnz + 1);
        flops = read_scalar(cum.ptr + F.nnz);
    }
    if (flops * 4 > A.nnz) return false;
    Y.clear();
    Y.nrows = F.nrows; Y.ncols = A.ncols; Y.W = W;
    Y.w.alloc(A.ncols * W);
    Y.w.zero();
    if (flops_out) *flops_out = flops;
    // a square operand with pull tables: write the result in its hot-set order, the form the next pull gathers from directly
    const u32 *perm = nullptr;
    if (lr && ctx().opt_perm_push && lr->perm_tag && lr->pperm && A.nrows == A.ncols && lr->pperm->n == A.ncols) {
        perm = lr->pperm->ptr;
        Y.pvert = lr->pvert;
        Y.perm_tag = lr->perm_tag;
    }
    if (flops == 0) return true;
    switch (W) {
    case 1: csr_push_impl<1>(F, A, Y, flops, erow.ptr, cum.ptr, perm); break;
    case 2: csr_push_impl<2>(F, A, Y, flops, erow.ptr, cum.ptr, perm); break;
    case 4: csr_push_impl<4>(F, A, Y, flops, erow.ptr, cum.ptr, perm); break;
    case 8: csr_push_impl<8>(F, A, Y, flops, erow.ptr, cum.ptr, perm); break;
    default: csr_push_impl<16>(F, A, Y, flops, erow.ptr, cum.ptr, perm); break;
    }
    return true;
}

// permuted form -> natural order: Xn[pvert[s]] = X[s], one thread per word
__global__ void k_unpermute(const u64 *__restrict__ X, const u32 *__restrict__ pvert, u64 n, u32 W, u64 *__restrict__ out) {
    u64 t = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    const u64 stride = (u64)gridDim.x * blockDim.x, total = n * W;
    for (; t < total; t += stride) { const u64 s = t / W, w = t - s * W; out[(u64)pvert[s] * W + w] = X[t]; }
}
void bits_naturalise(DevBits &X) {
    if (!X.permuted()) return;
    const u64 n = X.ncols;
    DevBuf<u64> out(n * X.W);
    if (n) LAUNCH(k_unpermute, grid_for(n * X.W, 256, 148 * 32), 256, 0, X.w.ptr, X.pvert->ptr, n, X.W, out.ptr);
    X.w = std::move(out);
    X.pvert.reset();
    X.perm_tag = 0;
}
static void require_natural(const DevBits &X, const char *what) {
    if (X.permuted()) throw GrbError(-101, std::string(what) + ": frontier is in permuted form (internal error: bits_naturalise was not called)");
}

// ---------------------------------------------------------------------------- pull
// 8 lanes per output vertex j: OR of X[k] over k in A'(j,:).  Rows longer than LONG_ROW are
// zeroed here and finished by k_bits_pull_long (several CTAs per row, RED.OR into Y).
struct OrOp64 { __device__ u64 operator()(u64 a, u64 b) const { return a | b; } };

// G[w] = OR over all vertices of word w: the terminal value of the OR monoid for this frontier.  A pull row whose
// accumulator has reached G cannot change any more, so the rest of its neighbours need not be gathered (exact).
template <int W>
__global__ void __launch_bounds__(256) k_or_all(const u64 *__restrict__ X, u64 nv, u64 *__restrict__ G) {
    typedef cub::BlockReduce<u64, 256> Red;
    __shared__ typename Red::TempStorage ts;
    u64 acc[W];
#pragma unroll
    for (int w = 0; w < W; w++) acc[w] = 0;
    u64 t = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    u64 stride = (u64)gridDim.x * blockDim.x;
    for (; t < nv; t += stride) {
#pragma unroll
        for (int w = 0; w < W; w++) acc[w] |= X[t * W + w];
    }
#pragma unroll
    for (int w = 0; w < W; w++) {
        u64 r = Red(ts).Reduce(acc[w], OrOp64());
        __syncthreads();
        if (threadIdx.x == 0 && r) atomicOr((unsigned long long *)&G[w], r);
    }
}

template <bool HINTS> __device__ __forceinline__ u32 ld_col(const u32 *p, u64 strm) { return HINTS ? ld_u32_stream(p, strm) : __ldg(p); }
template <bool HINTS> __device__ __forceinline__ u64 ld_ptr(const u64 *p, u64 strm) { return HINTS ? ld_u64_stream(p, strm) : __ldg(p); }
template <bool HINTS> __device__ __forceinline__ u64 ld_x(const u64 *p, u64 keep) { return HINTS ? ld_u64_hint(p, keep) : __ldg(p); }
// acc |= the W frontier words of one vertex: ONE 16-byte load per word pair (a W=4 vertex is a single 32 B sector), so the
// number of L1TEX wavefronts per gathered vertex does not grow with W
template <int W, bool HINTS> __device__ __forceinline__ void or_words(u64 (&acc)[W], const u64 *__restrict__ p, u64 keep) {
    if (W == 1) { acc[0] |= ld_x<HINTS>(p, keep); return; }
    if (W >= 4) {
#pragma unroll
        for (int w4 = 0; w4 < W / 4; w4++) {
            // plain 256-bit loads: a cache-hint operand is not encoded for LDG.256, and the .L2::evict_last / evict_first
            // qualifiers (hot prefix / cold tail) measured slower (1.72 vs 1.62 ms) -- with enough gathers in flight the
            // skewed access stream keeps its hot set in L2 by itself (scripts/ubench/gather_skew.cu)
            u64x4 v = ld_v4(p + 4 * w4);
            acc[4 * w4] |= v.a;
            acc[4 * w4 + 1] |= v.b;
            acc[4 * w4 + 2] |= v.c;
            acc[4 * w4 + 3] |= v.d;
        }
        return;
    }
    const ulonglong2 *q = reinterpret_cast<const ulonglong2 *>(p);
#pragma unroll
    for (int w2 = 0; w2 < W / 2; w2++) {
        ulonglong2 v = HINTS ? ld_v2_hint(q + w2, keep) : __ldg(q + w2);
        acc[2 * w2] |= v.x;
        acc[2 * w2 + 1] |= v.y;
    }
}

template <int W, bool HINTS, int U, bool EARLY>
__global__ void __launch_bounds__(256)
k_bits_pull(const u64 *__restrict__ ATp, const u32 *__restrict__ ATj, u64 n, const u64 *__restrict__ X,
            u64 *__restrict__ Y, const u64 *__restrict__ Gp, u32 hot_bytes, u32 tot_bytes) {
    const u32 lane8 = threadIdx.x & 7;
    const u32 sub = (threadIdx.x & 31) >> 3;
    const u32 gmask = 0xFFu << (8 * sub);            // the 8 lanes that share a row
    const u64 keep = HINTS ? policy_range(X, hot_bytes, tot_bytes) : 0, strm = policy_stream();
    u64 G[W];
#pragma unroll
    for (int w = 0; w < W; w++) G[w] = Gp ? Gp[w] : ~0ULL;
    u64 group = ((u64)blockIdx.x * blockDim.x + threadIdx.x) >> 3;
    u64 ngroups = ((u64)gridDim.x * blockDim.x) >> 3;
    u64 warp_first = group - sub; // group id of this warp's first 8-lane group
    // software pipeline: row pointers of the NEXT row are in flight while the current row gathers
    u64 ns = 0, ne = 0;
    if (warp_first + sub < n) { ns = ld_ptr<HINTS>(ATp + warp_first + sub, strm); ne = ld_ptr<HINTS>(ATp + warp_first + sub + 1, strm); }
    for (u64 base = warp_first; base < n; base += ngroups) {
        u64 j = base + sub;
        u64 s = ns, e = ne;
        u64 jn = j + ngroups;
        ns = 0; ne = 0;
        if (jn < n) { ns = ld_ptr<HINTS>(ATp + jn, strm); ne = ld_ptr<HINTS>(ATp + jn + 1, strm); }
        if (j >= n) { s = 0; e = 0; }
        if (e - s > LONG_ROW) e = s;
        u64 acc[W];
#pragma unroll
        for (int w = 0; w < W; w++) acc[w] = 0;
        if (EARLY) {
            // group-uniform trip count (all 8 lanes iterate together) so the lanes can vote inside the loop
            for (u64 qb = s; qb < e; qb += 8 * U) {
                u32 k[U];
#pragma unroll
                for (int u = 0; u < U; u++) { u64 q = qb + lane8 + 8 * u; k[u] = (q < e) ? ld_col<HINTS>(ATj + q, strm) : 0xFFFFFFFFu; }
#pragma unroll
                for (int u = 0; u < U; u++)
                    if (k[u] != 0xFFFFFFFFu) {
                        or_words<W, HINTS>(acc, X + (u64)k[u] * W, keep);
                    }
                if (qb + 8 * U < e) {   // more to come: stop if the row already holds the terminal value
                    // cheap test every iteration (each lane's own partial saturated), exact cross-lane OR every 4th
                    bool full = true;
                    if ((((qb - s) / (8 * U)) & 3) == 3) {
#pragma unroll
                        for (int w = 0; w < W; w++) {
                            u64 a = acc[w];
                            a |= __shfl_xor_sync(gmask, a, 1);
                            a |= __shfl_xor_sync(gmask, a, 2);
                            a |= __shfl_xor_sync(gmask, a, 4);
                            acc[w] = a;
                            full = full && (a == G[w]);
                        }
                    } else {
#pragma unroll
                        for (int w = 0; w < W; w++) full = full && (acc[w] == G[w]);
                        full = __all_sync(gmask, full);
                    }
                    if (full) break;
                }
            }
        } else {
            for (u64 q = s + lane8; q < e; q += 8 * U) {
                u32 k[U];
#pragma unroll
                for (int u = 0; u < U; u++) k[u] = (q + 8 * u < e) ? ld_col<HINTS>(ATj + q + 8 * u, strm) : 0xFFFFFFFFu;
#pragma unroll
                for (int u = 0; u < U; u++)
                    if (k[u] != 0xFFFFFFFFu) {
                        or_words<W, HINTS>(acc, X + (u64)k[u] * W, keep);
                    }
            }
        }
#pragma unroll
        for (int w = 0; w < W; w++) {
            u64 a = acc[w];
            a |= __shfl_xor_sync(0xffffffffu, a, 1);
            a |= __shfl_xor_sync(0xffffffffu, a, 2);
            a |= __shfl_xor_sync(0xffffffffu, a, 4);
            acc[w] = a;
        }
        if (j < n) {
#pragma unroll
            for (int w = 0; w < W; w++)
                if ((w & 7) == (int)lane8) { if (HINTS) st_u64_stream(Y + j * W + w, acc[w], strm); else Y[j * W + w] = acc[w]; }
        }
    }
}

// Software-pipelined 8-lanes-per-row pull (pull_kernel = 3, default).  The dependent chain of one row is
// rowptr -> col_idx batch -> X gathers; here all three stages of consecutive batches overlap: row pointers are loaded two
// rows ahead, and the col_idx batch that follows the current one (the next 8*U entries of the same row, or the first
// batch of the group's next row when this is the last) is in flight while the current batch gathers.
#ifndef PIPE_MINB
#define PIPE_MINB 0    // resident CTAs per SM the compiler must leave room for (0 = its own choice); fewer = more landing registers
#endif
#if PIPE_MINB > 0
#define PIPE_BOUNDS __launch_bounds__(256, PIPE_MINB)
#else
#define PIPE_BOUNDS __launch_bounds__(256)
#endif
template <int W, bool HINTS, int U, bool EARLY>
__global__ void PIPE_BOUNDS
k_bits_pull_pipe(const u64 *__restrict__ ATp, const u32 *__restrict__ ATj, u64 n, const u64 *__restrict__ X,
                 u64 *__restrict__ Y, const u64 *__restrict__ Gp, u32 hot_bytes, u32 tot_bytes) {
    const u32 lane8 = threadIdx.x & 7;
    const u32 sub = (threadIdx.x & 31) >> 3;
    const u32 gmask = 0xFFu << (8 * sub);
    const u64 keep = HINTS ? policy_range(X, hot_bytes, tot_bytes) : 0, strm = policy_stream();
    u64 G[W];
#pragma unroll
    for (int w = 0; w < W; w++) G[w] = (EARLY && Gp) ? Gp[w] : ~0ULL;
    const u64 ngroups = ((u64)gridDim.x * blockDim.x) >> 3;
    u64 j = ((u64)blockIdx.x * blockDim.x + threadIdx.x) >> 3;
    const u64 warp_first = j - sub;
    // rows j (s0,e0), j + ngroups (s1,e1); j + 2*ngroups is loaded inside the loop
    u64 s0 = 0, e0 = 0, s1 = 0, e1 = 0;
    if (j < n) { s0 = ld_ptr<HINTS>(ATp + j, strm); e0 = ld_ptr<HINTS>(ATp + j + 1, strm); }
    if (j + ngroups < n) { s1 = ld_ptr<HINTS>(ATp + j + ngroups, strm); e1 = ld_ptr<HINTS>(ATp + j + ngroups + 1, strm); }
    if (e0 - s0 > LONG_ROW) e0 = s0;
    u32 k[U];
#pragma unroll
    for (int u = 0; u < U; u++) { u64 q = s0 + lane8 + 8 * u; k[u] = (q < e0) ? ld_col<HINTS>(ATj + q, strm) : 0xFFFFFFFFu; }
    for (u64 base = warp_first; base < n; base += ngroups, j += ngroups) {
        u64 s2 = 0, e2 = 0;
        if (j + 2 * ngroups < n) { s2 = ld_ptr<HINTS>(ATp + j + 2 * ngroups, strm); e2 = ld_ptr<HINTS>(ATp + j + 2 * ngroups + 1, strm); }
        if (e1 - s1 > LONG_ROW) e1 = s1;
        u64 acc[W];
#pragma unroll
        for (int w = 0; w < W; w++) acc[w] = 0;
        u64 qb = s0;
        u32 it = 0;
        while (true) {
            const bool more = qb + 8 * U < e0;
            // col_idx of the following batch: same row if it continues, else the first batch of the next row
            const u64 nb = more ? qb + 8 * U : s1, ne = more ? e0 : e1;
            u32 kn[U];
#pragma unroll
            for (int u = 0; u < U; u++) { u64 q = nb + lane8 + 8 * u; kn[u] = (q < ne) ? ld_col<HINTS>(ATj + q, strm) : 0xFFFFFFFFu; }
#pragma unroll
            for (int u = 0; u < U; u++)
                if (k[u] != 0xFFFFFFFFu) or_words<W, HINTS>(acc, X + (u64)k[u] * W, keep);
#pragma unroll
            for (int u = 0; u < U; u++) k[u] = kn[u];
            if (!more) break;
            if (EARLY) {
                // cheap test every iteration (each lane holds a saturated partial), exact cross-lane OR every 4th
                bool full = true;
                if ((it & 3) == 3) {
#pragma unroll
                    for (int w = 0; w < W; w++) {
                        u64 a = acc[w];
                        a |= __shfl_xor_sync(gmask, a, 1);
                        a |= __shfl_xor_sync(gmask, a, 2);
                        a |= __shfl_xor_sync(gmask, a, 4);
                        acc[w] = a;
                        full = full && (a == G[w]);
                    }
                } else {
#pragma unroll
                    for (int w = 0; w < W; w++) full = full && (acc[w] == G[w]);
                    full = __all_sync(gmask, full);
                }
                if (full) {   // row saturated: the prefetched batch belongs to this row, fetch the next row's instead
#pragma unroll
                    for (int u = 0; u < U; u++) { u64 q = s1 + lane8 + 8 * u; k[u] = (q < e1) ? ld_col<HINTS>(ATj + q, strm) : 0xFFFFFFFFu; }
                    break;
                }
            }
            qb += 8 * U;
            it++;
        }
#pragma unroll
        for (int w = 0; w < W; w++) {
            u64 a = acc[w];
            a |= __shfl_xor_sync(gmask, a, 1);
            a |= __shfl_xor_sync(gmask, a, 2);
            a |= __shfl_xor_sync(gmask, a, 4);
            acc[w] = a;
        }
        if (j < n) {
#pragma unroll
            for (int w = 0; w < W; w++)
                if ((w & 7) == (int)lane8) { if (HINTS) st_u64_stream(Y + j * W + w, acc[w], strm); else Y[j * W + w] = acc[w]; }
        }
        s0 = s1; e0 = e1; s1 = s2; e1 = e2;
    }
}

// ---------------------------------------------------------------------------- degree-binned pull (pull_kernel = 4)
// The in-degree distribution of a power-law graph is what starves the 8-lane kernels: 86 % of the rows of A' hold <= 8
// entries (56 % none) and ride in the same warp as rows hundreds of entries long, so most 8-lane groups idle while one
// gathers -- and the rate of this kernel is set by the number of gathers in flight (scripts/ubench/gather_skew.cu: 240 G
// gathers/s with every lane busy, against ~70 G/s here).  Rows are therefore binned once per matrix:
//   small (<= SMALL_ROW) : one lane per row in natural order (k_bits_pull_small; also zero-fills empty and long rows)
//   mid   (<= LONG_ROW)  : a list sorted by length, descending; the four 8-lane groups of a warp take ADJACENT list entries,
//                          i.e. rows of (nearly) the same length, so they finish their batches together (k_bits_pull_mid)
//   long                 : k_bits_pull_long, as before
static const u32 SMALL_ROW = 8;

__global__ void k_flag_mid(const u64 *__restrict__ p, u64 n, u32 *__restrict__ flag) {
    u64 r = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    u64 stride = (u64)gridDim.x * blockDim.x;
    for (; r < n; r += stride) {
        u64 d = p[r + 1] - p[r];
        flag[r] = (d > SMALL_ROW && d <= LONG_ROW) ? 1u : 0u;
    }
}
// key = (LONG_ROW - len) << 32 | row : ascending sort = length descending, row ascending inside one length
__global__ void k_mid_keys(const u32 *__restrict__ flag, const u64 *__restrict__ pos, u64 n, const u64 *__restrict__ p,
                           u64 *__restrict__ keys) {
    u64 r = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    u64 stride = (u64)gridDim.x * blockDim.x;
    for (; r < n; r += stride)
        if (flag[r]) keys[pos[r]] = ((LONG_ROW - (p[r + 1] - p[r])) << 32) | r;
}
__global__ void k_mid_unpack(const u64 *__restrict__ keys, u64 nm, const u64 *__restrict__ p, u32 *__restrict__ m_row,
                             u32 *__restrict__ m_len, u64 *__restrict__ m_start) {
    u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nm) return;
    u32 r = (u32)keys[i];
    m_row[i] = r;
    m_start[i] = p[r];
    m_len[i] = (u32)(p[r + 1] - p[r]);
}
static void build_mid_list(const DevCSR &AT, LongRows &lr) {
    u64 n = AT.nrows;
    lr.nm = 0;
    lr.m_row.release(); lr.m_len.release(); lr.m_start.release();
    if (!n) return;
    DevBuf<u32> flag(n + 1);
    DevBuf<u64> pos(n + 1);
    LAUNCH(k_flag_mid, grid_for(n, 256, 1 << 16), 256, 0, AT.p.ptr, n, flag.ptr);
    CUDA_TRY(cudaMemsetAsync(flag.ptr + n, 0, sizeof(u32), stream()));
    exclusive_scan_u32_to_u64(flag.ptr, pos.ptr, n + 1);
    u64 nm = read_scalar(pos.ptr + n);
    lr.nm = nm;
    if (!nm) return;
    DevBuf<u64> keys(nm);
    LAUNCH(k_mid_keys, grid_for(n, 256, 1 << 16), 256, 0, flag.ptr, pos.ptr, n, AT.p.ptr, keys.ptr);
    sort_keys_u64(keys.ptr, nm, 46);            // LONG_ROW = 2^12: 32 row bits + 13 length bits, rounded up
    lr.m_row.alloc(nm); lr.m_len.alloc(nm); lr.m_start.alloc(nm);
    LAUNCH(k_mid_unpack, grid_for(nm, 256), 256, 0, keys.ptr, nm, AT.p.ptr, lr.m_row.ptr, lr.m_len.ptr, lr.m_start.ptr);
}

template <int W> __device__ __forceinline__ void store_row(u64 *dst, const u64 (&acc)[W]) {
    if constexpr (W == 1) { dst[0] = acc[0]; }
    else if constexpr (W == 2) { *reinterpret_cast<ulonglong2 *>(dst) = make_ulonglong2(acc[0], acc[1]); }
    else {
#pragma unroll
        for (int w4 = 0; w4 < W / 4; w4++)
            asm volatile("st.global.v4.u64 [%0], {%1, %2, %3, %4};" ::"l"(dst + 4 * w4), "l"(acc[4 * w4]), "l"(acc[4 * w4 + 1]),
                         "l"(acc[4 * w4 + 2]), "l"(acc[4 * w4 + 3]) : "memory");
    }
}

// one lane per row, natural order: rows of <= SMALL_ROW entries are finished here, long rows are zero-filled (k_bits_pull_long
// ORs into them), mid rows are left to k_bits_pull_mid
template <int W, bool HINTS>
__global__ void __launch_bounds__(256)
k_bits_pull_small(const u64 *__restrict__ ATp, const u32 *__restrict__ ATj, u64 n, const u64 *__restrict__ X,
                  u64 *__restrict__ Y, u32 hot_bytes, u32 tot_bytes) {
    const u64 keep = HINTS ? policy_range(X, hot_bytes, tot_bytes) : 0;
    u64 j = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    const u64 stride = (u64)gridDim.x * blockDim.x;
    // three-stage pipeline per lane: row pointers two rows ahead, first col_idx batch one row ahead, gathers now
    u64 s = 0, e = 0, s1 = 0, e1 = 0;
    if (j < n) { s = __ldg(ATp + j); e = __ldg(ATp + j + 1); }
    if (j + stride < n) { s1 = __ldg(ATp + j + stride); e1 = __ldg(ATp + j + stride + 1); }
    u32 k[4];
#pragma unroll
    for (int u = 0; u < 4; u++) k[u] = (e - s <= SMALL_ROW && s + u < e) ? __ldg(ATj + s + u) : 0xFFFFFFFFu;
    for (; j < n; j += stride) {
        u64 s2 = 0, e2 = 0;
        if (j + 2 * stride < n) { s2 = __ldg(ATp + j + 2 * stride); e2 = __ldg(ATp + j + 2 * stride + 1); }
        u32 kn[4];
#pragma unroll
        for (int u = 0; u < 4; u++) kn[u] = (e1 - s1 <= SMALL_ROW && s1 + u < e1) ? __ldg(ATj + s1 + u) : 0xFFFFFFFFu;
        u64 acc[W];
#pragma unroll
        for (int w = 0; w < W; w++) acc[w] = 0;
        if (e - s <= SMALL_ROW) {
#pragma unroll
            for (int u = 0; u < 4; u++)
                if (k[u] != 0xFFFFFFFFu) or_words<W, HINTS>(acc, X + (u64)k[u] * W, keep);
            if (e - s > 4) {
#pragma unroll
                for (int u = 0; u < 4; u++) k[u] = (s + 4 + u < e) ? __ldg(ATj + s + 4 + u) : 0xFFFFFFFFu;
#pragma unroll
                for (int u = 0; u < 4; u++)
                    if (k[u] != 0xFFFFFFFFu) or_words<W, HINTS>(acc, X + (u64)k[u] * W, keep);
            }
        }
        if (e - s <= SMALL_ROW || e - s > LONG_ROW) store_row<W>(Y + j * W, acc);
#pragma unroll
        for (int u = 0; u < 4; u++) k[u] = kn[u];
        s = s1; e = e1; s1 = s2; e1 = e2;
    }
}

// 8 lanes per mid row, rows taken from the length-sorted list; same three-stage software pipeline as k_bits_pull_pipe
// (list entry two steps ahead, col_idx batch one step ahead, gathers now).  Step t of group g handles list entry
// t * ngroups + g on even steps and t * ngroups + (ngroups - 1 - g) on odd ones, which evens out the totals per group.
template <int W, bool HINTS, int U, bool EARLY>
__global__ void PIPE_BOUNDS
k_bits_pull_mid(const u32 *__restrict__ m_row, const u64 *__restrict__ m_start, const u32 *__restrict__ m_len, u64 nm,
                const u32 *__restrict__ ATj, const u64 *__restrict__ X, u64 *__restrict__ Y, const u64 *__restrict__ Gp,
                u32 hot_bytes, u32 tot_bytes) {
    const u32 lane8 = threadIdx.x & 7;
    const u32 sub = (threadIdx.x & 31) >> 3;
    const u32 gmask = 0xFFu << (8 * sub);
    const u64 keep = HINTS ? policy_range(X, hot_bytes, tot_bytes) : 0, strm = policy_stream();
    u64 G[W];
#pragma unroll
    for (int w = 0; w < W; w++) G[w] = (EARLY && Gp) ? Gp[w] : ~0ULL;
    const u64 ngroups = ((u64)gridDim.x * blockDim.x) >> 3;
    const u64 g = ((u64)blockIdx.x * blockDim.x + threadIdx.x) >> 3;
    const u64 steps = (nm + ngroups - 1) / ngroups;
    auto entry = [&](u64 t) -> u64 { return t * ngroups + ((t & 1) ? ngroups - 1 - g : g); };
    auto fetch = [&](u64 t, u64 &s, u64 &e) {
        s = 0; e = 0;
        if (t < steps) { u64 i = entry(t); if (i < nm) { s = m_start[i]; e = s + m_len[i]; } }
    };
    u64 s0, e0, s1, e1;
    fetch(0, s0, e0);
    fetch(1, s1, e1);
    u32 k[U];
#pragma unroll
    for (int u = 0; u < U; u++) { u64 q = s0 + lane8 + 8 * u; k[u] = (q < e0) ? ld_col<HINTS>(ATj + q, strm) : 0xFFFFFFFFu; }
    for (u64 t = 0; t < steps; t++) {
        u64 s2, e2;
        fetch(t + 2, s2, e2);
        u64 acc[W];
#pragma unroll
        for (int w = 0; w < W; w++) acc[w] = 0;
        u64 qb = s0;
        u32 it = 0;
        while (true) {
            const bool more = qb + 8 * U < e0;
            const u64 nb = more ? qb + 8 * U : s1, ne = more ? e0 : e1;
            u32 kn[U];
#pragma unroll
            for (int u = 0; u < U; u++) { u64 q = nb + lane8 + 8 * u; kn[u] = (q < ne) ? ld_col<HINTS>(ATj + q, strm) : 0xFFFFFFFFu; }
#pragma unroll
            for (int u = 0; u < U; u++)
                if (k[u] != 0xFFFFFFFFu) or_words<W, HINTS>(acc, X + (u64)k[u] * W, keep);
#pragma unroll
            for (int u = 0; u < U; u++) k[u] = kn[u];
            if (!more) break;
            if (EARLY) {
                bool full = true;
                if ((it & 3) == 3) {
#pragma unroll
                    for (int w = 0; w < W; w++) {
                        u64 a = acc[w];
                        a |= __shfl_xor_sync(gmask, a, 1);
                        a |= __shfl_xor_sync(gmask, a, 2);
                        a |= __shfl_xor_sync(gmask, a, 4);
                        acc[w] = a;
                        full = full && (a == G[w]);
                    }
                } else {
#pragma unroll
                    for (int w = 0; w < W; w++) full = full && (acc[w] == G[w]);
                    full = __all_sync(gmask, full);
                }
                if (full) {
#pragma unroll
                    for (int u = 0; u < U; u++) { u64 q = s1 + lane8 + 8 * u; k[u] = (q < e1) ? ld_col<HINTS>(ATj + q, strm) : 0xFFFFFFFFu; }
                    break;
                }
            }
            qb += 8 * U;
            it++;
        }
#pragma unroll
        for (int w = 0; w < W; w++) {
            u64 a = acc[w];
            a |= __shfl_xor_sync(gmask, a, 1);
            a |= __shfl_xor_sync(gmask, a, 2);
            a |= __shfl_xor_sync(gmask, a, 4);
            acc[w] = a;
        }
        if (e0 > s0 && lane8 == 0) {
            const u64 i = entry(t);
            store_row<W>(Y + (u64)m_row[i] * W, acc);
        }
        s0 = s1; e0 = e1; s1 = s2; e1 = e2;
    }
}

template <int W>
__global__ void __launch_bounds__(256)
k_bits_pull_long(const u32 *__restrict__ lrows, const u64 *__restrict__ choff, u32 nlong, const u64 *__restrict__ ATp,
                 const u32 *__restrict__ ATj, const u64 *__restrict__ X, u64 *__restrict__ Y, const u64 *__restrict__ Gp) {
    typedef cub::BlockReduce<u64, 256> Red;
    __shared__ typename Red::TempStorage ts;
    __shared__ u32 s_li;
    if (threadIdx.x == 0) {   // chunk -> (long row, chunk within the row): binary search in the per-row chunk prefix
        u64 c = blockIdx.x, lo = 0, hi = nlong - 1;
        while (lo < hi) { u64 mid = (lo + hi + 1) >> 1; if (choff[mid] <= c) lo = mid; else hi = mid - 1; }
        s_li = (u32)lo;
    }
    __syncthreads();
    const u32 li = s_li;
    u32 j = lrows[li];
    u64 s = ATp[j], e = ATp[j + 1];
    u64 c0 = s + ((u64)blockIdx.x - choff[li]) * LONG_CHUNK;
    if (c0 >= e) return;
    u64 c1 = c0 + LONG_CHUNK;
    if (c1 > e) c1 = e;
    u64 acc[W];
#pragma unroll
    for (int w = 0; w < W; w++) acc[w] = 0;
    u64 G[W];
#pragma unroll
    for (int w = 0; w < W; w++) G[w] = Gp ? Gp[w] : ~0ULL;
    // warp-uniform loop; leaves as soon as the warp holds the terminal value: cheap test every iteration (each lane holds a
    // saturated partial), exact cross-lane OR every 4th
    u32 it = 0;
    for (u64 qb = c0 + (threadIdx.x & ~31u); qb < c1; qb += 256, it++) {
        u64 q = qb + (threadIdx.x & 31);
        if (q < c1) {
            u32 k = __ldg(ATj + q);
            or_words<W, false>(acc, X + (u64)k * W, 0);
        }
        bool full = true;
        if ((it & 3) == 3) {
#pragma unroll
            for (int w = 0; w < W; w++) {
                u64 a = acc[w];
#pragma unroll
                for (int d = 16; d; d >>= 1) a |= __shfl_xor_sync(0xffffffffu, a, d);
                acc[w] = a;
                full = full && (a == G[w]);
            }
        } else {
#pragma unroll
            for (int w = 0; w < W; w++) full = full && (acc[w] == G[w]);
            full = __all_sync(0xffffffffu, full);
        }
        if (full) break;
    }
#pragma unroll
    for (int w = 0; w < W; w++) {
        u64 r = Red(ts).Reduce(acc[w], OrOp64());
        __syncthreads();
        if (threadIdx.x == 0 && r) atomicOr((unsigned long long *)&Y[(u64)j * W + w], r);
    }
}

// ---------------------------------------------------------------------------- lane-split pull (pull_kernel = 5, default)
// What bounds a pull is the L1TEX tag stage: one 128-byte line per clock per SM (scripts/ubench/gather*.cu: 281 G lines/s on
// 148 SMs), and a warp-wide gather with every lane on a different vertex costs 32 of those "wavefronts" per instruction.  A
// vertex record of W = 8 words is 64 bytes -- one line -- but a lane can fetch at most 32 bytes per instruction (LDG.256),
// so one-lane-per-vertex spends TWO wavefronts per vertex.  Here SPLIT = W / 4 adjacent lanes share a vertex: lane h of the
// pair loads words [4h, 4h+4) of the record with one LDG.256, the two requests fall into the same line and coalesce into one
// wavefront, and the pair never exchanges data (each lane ORs, keeps and finally stores its own 32-byte part of the output row).
// W = 8 -> half the wavefronts per gathered vertex; W = 16 -> a quarter; W <= 4 is unchanged.
// Rows longer than LONG_ROW are no longer a separate kernel: they enter the length-sorted list as LONG_ROW-entry segments whose
// partial results are merged with RED.OR (the small-row kernel zero-fills those rows first).
template <int WL> __device__ __forceinline__ void store_part(u64 *dst, const u64 (&acc)[WL]) { store_row<WL>(dst, acc); }

static const u32 SEG_ATOMIC = 0x80000000u;

// segments per row: 0 (small), 1 (mid), ceil(len / LONG_ROW) (long)
__global__ void k_seg_count(const u64 *__restrict__ p, u64 n, u32 *__restrict__ cnt) {
    u64 r = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    u64 stride = (u64)gridDim.x * blockDim.x;
    for (; r <= n; r += stride) {
        if (r == n) { cnt[r] = 0; break; }
        u64 d = p[r + 1] - p[r];
        cnt[r] = d <= SMALL_ROW ? 0u : (u32)((d + LONG_ROW - 1) / LONG_ROW);
    }
}
// key = LONG_ROW - seglen (ascending sort = longest first), value = position of the segment: row << 20 | index inside the row
__global__ void k_seg_emit(const u64 *__restrict__ p, u64 n, const u64 *__restrict__ pos, u64 *__restrict__ keys, u64 *__restrict__ vals) {
    u64 r = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    u64 stride = (u64)gridDim.x * blockDim.x;
    for (; r < n; r += stride) {
        const u64 d = p[r + 1] - p[r];
        if (d <= SMALL_ROW) continue;
        u64 o = pos[r];
        for (u64 s = 0, i = 0; s < d; s += LONG_ROW, i++) {
            const u64 len = d - s < LONG_ROW ? d - s : LONG_ROW;
            keys[o] = LONG_ROW - len;
            vals[o] = (r << 20) | i;
            o++;
        }
    }
}
__global__ void k_seg_unpack(const u64 *__restrict__ vals, u64 ns, const u64 *__restrict__ p, u32 *__restrict__ s_row,
                             u32 *__restrict__ s_len, u64 *__restrict__ s_start) {
    u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= ns) return;
    const u64 v = vals[i];
    const u32 r = (u32)(v >> 20);
    const u64 k = v & 0xFFFFFu;
    const u64 d = p[r + 1] - p[r];
    const u64 s = k * LONG_ROW;
    const u64 len = d - s < LONG_ROW ? d - s : LONG_ROW;
    s_row[i] = r;
    s_start[i] = p[r] + s;
    s_len[i] = (u32)len | (d > LONG_ROW ? SEG_ATOMIC : 0u);
}
static void build_seg_list(const DevCSR &AT, LongRows &lr) {
    const u64 n = AT.nrows;
    lr.ns = 0;
    lr.s_row.release(); lr.s_len.release(); lr.s_start.release();
    lr.seg_built = true;
    if (!n) return;
    if (lr.maxdeg >= ((u64)LONG_ROW << 20)) throw GrbError(-8, "pull: a row of A' exceeds 2^32 entries");
    DevBuf<u32> cnt(n + 1);
    DevBuf<u64> pos(n + 1);
    LAUNCH(k_seg_count, grid_for(n + 1, 256, 1 << 16), 256, 0, AT.p.ptr, n, cnt.ptr);
    exclusive_scan_u32_to_u64(cnt.ptr, pos.ptr, n + 1);
    const u64 ns = read_scalar(pos.ptr + n);
    lr.ns = ns;
    if (!ns) return;
    DevBuf<u64> keys(ns), vals(ns);
    LAUNCH(k_seg_emit, grid_for(n, 256, 1 << 16), 256, 0, AT.p.ptr, n, pos.ptr, keys.ptr, vals.ptr);
    sort_pairs_u64(keys.ptr, vals.ptr, ns, 13);          // stable: rows ascending inside one length
    lr.s_row.alloc(ns); lr.s_len.alloc(ns); lr.s_start.alloc(ns);
    LAUNCH(k_seg_unpack, grid_for(ns, 256), 256, 0, vals.ptr, ns, AT.p.ptr, lr.s_row.ptr, lr.s_len.ptr, lr.s_start.ptr);
}

// SPLIT lanes per row, natural order: rows of <= SMALL_ROW entries are finished here, rows longer than LONG_ROW are zero-filled
// (their segments arrive through RED.OR), everything in between is left to k_pull_seg.  This kernel is latency-bound, not
// L1TEX-bound (l1tex 17 %, issue 17 %: profiles/r2a): splitting the record across lanes halves the bytes each warp keeps in
// flight and was slower (0.90 vs 0.52 ms at W = 8), so by default one lane owns the whole record here (small_split = 0).
template <int W, bool HINTS, int SPLIT>
__global__ void __launch_bounds__(256)
k_pull_small(const u64 *__restrict__ ATp, const u32 *__restrict__ ATj, u64 n, const u64 *__restrict__ X,
             u64 *__restrict__ Y, u32 hot_bytes, u32 tot_bytes) {
    constexpr int WL = W / SPLIT;                  // words per lane: SPLIT = PullCfg<W>::SPLIT (one 32-byte load per record part) or 1
    const u64 keep = HINTS ? policy_range(X, hot_bytes, tot_bytes) : 0;
    const u64 t = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    const u32 h = (u32)(t % SPLIT);
    const u64 *Xh = X + h * WL;
    u64 j = t / SPLIT;
    const u64 stride = ((u64)gridDim.x * blockDim.x) / SPLIT;
    u64 s = 0, e = 0, s1 = 0, e1 = 0;
    if (j < n) { s = __ldg(ATp + j); e = __ldg(ATp + j + 1); }
    if (j + stride < n) { s1 = __ldg(ATp + j + stride); e1 = __ldg(ATp + j + stride + 1); }
    u32 k[4];
#pragma unroll
    for (int u = 0; u < 4; u++) k[u] = (e - s <= SMALL_ROW && s + u < e) ? __ldg(ATj + s + u) : 0xFFFFFFFFu;
    for (; j < n; j += stride) {
        u64 s2 = 0, e2 = 0;
        if (j + 2 * stride < n) { s2 = __ldg(ATp + j + 2 * stride); e2 = __ldg(ATp + j + 2 * stride + 1); }
        u32 kn[4];
#pragma unroll
        for (int u = 0; u < 4; u++) kn[u] = (e1 - s1 <= SMALL_ROW && s1 + u < e1) ? __ldg(ATj + s1 + u) : 0xFFFFFFFFu;
        u64 acc[WL];
#pragma unroll
        for (int w = 0; w < WL; w++) acc[w] = 0;
        if (e - s <= SMALL_ROW) {
#pragma unroll
            for (int u = 0; u < 4; u++)
                if (k[u] != 0xFFFFFFFFu) or_words<WL, HINTS>(acc, Xh + (u64)k[u] * W, keep);
            if (e - s > 4) {
#pragma unroll
                for (int u = 0; u < 4; u++) k[u] = (s + 4 + u < e) ? __ldg(ATj + s + 4 + u) : 0xFFFFFFFFu;
#pragma unroll
                for (int u = 0; u < 4; u++)
                    if (k[u] != 0xFFFFFFFFu) or_words<WL, HINTS>(acc, Xh + (u64)k[u] * W, keep);
            }
        }
        if (e - s <= SMALL_ROW || e - s > LONG_ROW) store_part<WL>(Y + j * W + h * WL, acc);
#pragma unroll
        for (int u = 0; u < 4; u++) k[u] = kn[u];
        s = s1; e = e1; s1 = s2; e1 = e2;
    }
}

// 8 lanes per list entry (a mid row, or one LONG_ROW-entry segment of a long row); entries sorted by length, descending.
// Same three-stage software pipeline as k_bits_pull_mid (list entry two steps ahead, col_idx batch one step ahead, gathers now);
// the 8 lanes cover VG vertices x SPLIT record parts per unroll step.
template <int W, bool HINTS, int U, bool EARLY>
__global__ void PIPE_BOUNDS
k_pull_seg(const u32 *__restrict__ s_row, const u64 *__restrict__ s_start, const u32 *__restrict__ s_len, u64 ns,
           const u32 *__restrict__ ATj, const u64 *__restrict__ X, u64 *__restrict__ Y, const u64 *__restrict__ Gp,
           u32 hot_bytes, u32 tot_bytes) {
    constexpr int WL = PullCfg<W>::WL, SPLIT = PullCfg<W>::SPLIT, VG = PullCfg<W>::VG;
    const u32 lane8 = threadIdx.x & 7;
    const u32 sub = (threadIdx.x & 31) >> 3;
    const u32 gmask = 0xFFu << (8 * sub);
    const u32 vi = lane8 / SPLIT, h = lane8 % SPLIT;
    const u64 *Xh = X + h * WL;
    const u64 keep = HINTS ? policy_range(X, hot_bytes, tot_bytes) : 0, strm = policy_stream();
    u64 G[WL];
#pragma unroll
    for (int w = 0; w < WL; w++) G[w] = (EARLY && Gp) ? Gp[h * WL + w] : ~0ULL;
    const u64 ngroups = ((u64)gridDim.x * blockDim.x) >> 3;
    const u64 g = ((u64)blockIdx.x * blockDim.x + threadIdx.x) >> 3;
    const u64 steps = (ns + ngroups - 1) / ngroups;
    auto entry = [&](u64 t) -> u64 { return t * ngroups + ((t & 1) ? ngroups - 1 - g : g); };
    auto fetch = [&](u64 t, u64 &s, u64 &e) {
        s = 0; e = 0;
        if (t < steps) { u64 i = entry(t); if (i < ns) { s = s_start[i]; e = s + (s_len[i] & ~SEG_ATOMIC); } }
    };
    u64 s0, e0, s1, e1;
    fetch(0, s0, e0);
    fetch(1, s1, e1);
    u32 k[U];
#pragma unroll
    for (int u = 0; u < U; u++) { u64 q = s0 + vi + VG * u; k[u] = (q < e0) ? ld_col<HINTS>(ATj + q, strm) : 0xFFFFFFFFu; }
    for (u64 t = 0; t < steps; t++) {
        u64 s2, e2;
        fetch(t + 2, s2, e2);
        u64 acc[WL];
#pragma unroll
        for (int w = 0; w < WL; w++) acc[w] = 0;
        u64 qb = s0;
        u32 it = 0;
        while (true) {
            const bool more = qb + VG * U < e0;
            const u64 nb = more ? qb + VG * U : s1, ne = more ? e0 : e1;
            u32 kn[U];
#pragma unroll
            for (int u = 0; u < U; u++) { u64 q = nb + vi + VG * u; kn[u] = (q < ne) ? ld_col<HINTS>(ATj + q, strm) : 0xFFFFFFFFu; }
#pragma unroll
            for (int u = 0; u < U; u++)
                if (k[u] != 0xFFFFFFFFu) or_words<WL, HINTS>(acc, Xh + (u64)k[u] * W, keep);
#pragma unroll
            for (int u = 0; u < U; u++) k[u] = kn[u];
            if (!more) break;
            if (EARLY) {   // stop once the entry holds the OR monoid's terminal value: cheap per-lane test, exact cross-lane OR every 4th
                bool full = true;
                if ((it & 3) == 3) {
#pragma unroll
                    for (int w = 0; w < WL; w++) {
                        u64 a = acc[w];
#pragma unroll
                        for (int d = SPLIT; d < 8; d <<= 1) a |= __shfl_xor_sync(gmask, a, d);
                        acc[w] = a;
                        full = full && (a == G[w]);
                    }
                } else {
#pragma unroll
                    for (int w = 0; w < WL; w++) full = full && (acc[w] == G[w]);
                }
                full = __all_sync(gmask, full);
                if (full) {
#pragma unroll
                    for (int u = 0; u < U; u++) { u64 q = s1 + vi + VG * u; k[u] = (q < e1) ? ld_col<HINTS>(ATj + q, strm) : 0xFFFFFFFFu; }
                    break;
                }
            }
            qb += VG * U;
            it++;
        }
#pragma unroll
        for (int w = 0; w < WL; w++) {
            u64 a = acc[w];
#pragma unroll
            for (int d = SPLIT; d < 8; d <<= 1) a |= __shfl_xor_sync(gmask, a, d);
            acc[w] = a;
        }
        if (e0 > s0 && vi == 0) {
            const u64 i = entry(t);
            u64 *dst = Y + (u64)s_row[i] * W + h * WL;
            if (s_len[i] & SEG_ATOMIC) {
#pragma unroll
                for (int w = 0; w < WL; w++) if (acc[w]) atomicOr((unsigned long long *)(dst + w), acc[w]);
            } else store_part<WL>(dst, acc);
        }
        s0 = s1; e0 = e1; s1 = s2; e1 = e2;
    }
}

// ---------------------------------------------------------------------------- pull, merge-path variant
// Balanced over (rows + nnz) of A' exactly like merge-based CSR SpMV: every CTA takes TILE consecutive items of the
// merged (row-end, nnz) sequence, so hub rows and runs of empty rows cost what they weigh.  Phase 1 gathers X[col]
// for the tile's nnz with a flat coalesced mapping (IPT independent gather chains per thread: memory-level
// parallelism), phase 2 is each thread's serial walk over its IPT merged items out of shared memory, phase 3 writes
// finished rows with plain stores (tile-boundary rows with RED.OR; Y is pre-zeroed).
template <int W> struct MpCfg { static const int TILE = (W <= 2) ? 2048 : (W == 4 ? 1024 : (W == 8 ? 512 : 256)); };

template <int W>
__global__ void __launch_bounds__(256)
k_bits_pull_mp(const u64 *__restrict__ ATp, const u32 *__restrict__ ATj, u64 n, u64 nnz, const u64 *__restrict__ X,
               u64 *__restrict__ Y, const u64 *__restrict__ mp_r /* row coordinate of every 256th diagonal */) {
    constexpr int T = 256, TILE = MpCfg<W>::TILE, IPT = TILE / T;
    extern __shared__ u64 smem_mp[];
    u64 *sx = smem_mp;                         // [TILE * W]   gathered words, nnz order
    u64 *rowacc = sx + (size_t)TILE * W;       // [(TILE + 1) * W]
    u32 *rend = (u32 *)(rowacc + (size_t)(TILE + 1) * W); // [TILE + 2] row ends relative to the tile's first nnz
    const u32 tid = threadIdx.x;
    const u64 keep = policy_keep(), strm = policy_stream();
    const u64 total = n + nnz;
    u64 d0 = (u64)blockIdx.x * TILE, d1 = d0 + TILE;
    if (d1 > total) d1 = total;
    // tile coordinates come from the per-matrix table built once by k_mp_coords (static for an immutable base)
    const u64 r0 = mp_r[d0 / 256], r1 = (d1 == total) ? n : mp_r[d1 / 256];
    const u64 z0 = d0 - r0, z1 = d1 - r1;
    const u32 nr = (u32)(r1 - r0);   // row-ends consumed in this tile; row r1 may be entered but not finished
    const u32 nz = (u32)(z1 - z0);
    for (u32 i = tid; i <= nr; i += T) {
        u64 r = r0 + i;
        u64 e = (r < n) ? ld_u64_stream(ATp + r + 1, strm) : nnz;
        u64 rel = e - z0;
        rend[i] = rel > 0xffffffffULL ? 0xffffffffu : (u32)rel;  // rows ending past the tile clamp high
#pragma unroll
        for (int w = 0; w < W; w++) rowacc[(size_t)i * W + w] = 0;
    }
#pragma unroll
    for (int k = 0; k < IPT; k++) {
        u32 e = tid + k * T;
        if (e < nz) {
            u32 col = ld_u32_stream(ATj + z0 + e, strm);
#pragma unroll
            for (int w = 0; w < W; w++) sx[(size_t)e * W + w] = ld_u64_hint(X + (u64)col * W + w, keep);
        }
    }
    __syncthreads();
    // per-thread merge-path start inside the tile
    u32 dt = tid * IPT;
    u32 items = (u32)(d1 - d0);
    if (dt > items) dt = items;
    u32 lo = dt > nz ? dt - nz : 0, hi = dt < nr ? dt : nr;
    while (lo < hi) {
        u32 mid = (lo + hi) >> 1;
        if (rend[mid] <= dt - mid - 1) lo = mid + 1; else hi = mid;
    }
    u32 r = lo, z = dt - lo;
    u32 dend = dt + IPT;
    if (dend > items) dend = items;
    u64 acc[W];
#pragma unroll
    for (int w = 0; w < W; w++) acc[w] = 0;
    auto flush = [&](u32 row) {
#pragma unroll
        for (int w = 0; w < W; w++) {
            u64 a = acc[w];
            if (a) {
                u32 *p32 = (u32 *)&rowacc[(size_t)row * W + w];
                if ((u32)a) atomicOr(p32, (u32)a);
                if ((u32)(a >> 32)) atomicOr(p32 + 1, (u32)(a >> 32));
            }
            acc[w] = 0;
        }
    };
    for (u32 d = dt; d < dend; d++) {
        if (r < nr ? (z < rend[r]) : true) {          // consume one nnz of row r (r == nr: the unfinished last row)
#pragma unroll
            for (int w = 0; w < W; w++) acc[w] |= sx[(size_t)z * W + w];
            z++;
        } else {
            flush(r);
            r++;
        }
    }
    flush(r);
    __syncthreads();
    for (u32 i = tid; i <= nr; i += T) {
        u64 row = r0 + i;
        if (row >= n) continue;
#pragma unroll
        for (int w = 0; w < W; w++) {
            u64 v = rowacc[(size_t)i * W + w];
            if (v) {
                if (i == 0 || i == nr) atomicOr((unsigned long long *)&Y[row * W + w], v);
                else st_u64_stream(Y + row * W + w, v, strm);
            }
        }
    }
}

// merge-path coordinates of every 256th diagonal of (row-ends, nnz): r = #rows finished before the diagonal
__global__ void k_mp_coords(const u64 *__restrict__ ATp, u64 n, u64 nnz, u64 ndiag, u64 *__restrict__ mp_r) {
    u64 t = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= ndiag) return;
    u64 d = t * 256;
    u64 total = n + nnz;
    if (d > total) d = total;
    u64 lo = d > nnz ? d - nnz : 0, hi = d < n ? d : n;
    while (lo < hi) {
        u64 mid = (lo + hi) >> 1;
        if (ATp[mid + 1] <= d - mid - 1) lo = mid + 1; else hi = mid;
    }
    mp_r[t] = lo;
}

__global__ void k_flag_long(const u64 *__restrict__ p, u64 n, u32 *__restrict__ flag, u64 *__restrict__ maxdeg) {
    u64 r = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    u64 stride = (u64)gridDim.x * blockDim.x;
    u64 mx = 0;
    for (; r < n; r += stride) {
        u64 d = p[r + 1] - p[r];
        flag[r] = d > LONG_ROW ? 1u : 0u;
        if (d > mx) mx = d;
    }
    if (mx > LONG_ROW) atomicMax((unsigned long long *)maxdeg, mx);
}
__global__ void k_long_chunks(const u32 *__restrict__ lrows, u64 nl, const u64 *__restrict__ p, u64 *__restrict__ nch) {
    u64 t = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (t > nl) return;
    if (t == nl) { nch[t] = 0; return; }
    u32 r = lrows[t];
    nch[t] = (p[r + 1] - p[r] + LONG_CHUNK - 1) / LONG_CHUNK;
}
__global__ void k_scatter_flagged(const u32 *__restrict__ flag, const u64 *__restrict__ pos, u64 n, u32 *__restrict__ out) {
    u64 r = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    u64 stride = (u64)gridDim.x * blockDim.x;
    for (; r < n; r += stride)
        if (flag[r]) out[pos[r]] = (u32)r;
}

// ---- hot-set packing of the gather index space ------------------------------------------------------------------
// The pull gathers X[k] with k distributed like the out-degree of k.  Relabel k -> slot, slots ordered by out-degree
// descending and restricted to vertices that have out-edges at all (sinks are never gathered): the hot part of the
// frontier words becomes contiguous (L1/L2 resident), and the gather footprint shrinks from 8*W*n to 8*W*n1 bytes.
__global__ void k_degree_keys(const u64 *__restrict__ Ap, u64 n, u64 *__restrict__ keys, u64 *__restrict__ n1) {
    u64 v = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    u64 stride = (u64)gridDim.x * blockDim.x;
    u64 cnt = 0;
    for (; v < n; v += stride) {
        u64 d = Ap[v + 1] - Ap[v];
        if (d > 0xFFFFFFFEULL) d = 0xFFFFFFFEULL;
        keys[v] = ((0xFFFFFFFFULL - d) << 32) | v;   // ascending sort => degree descending, then vertex id ascending
        cnt += d != 0;
    }
    if (cnt) atomicAdd((unsigned long long *)n1, cnt);
}
__global__ void k_slots_from_keys(const u64 *__restrict__ keys, u64 n, u64 n1, u32 *__restrict__ vert, u32 *__restrict__ slot,
                                  u32 *__restrict__ hdeg) {
    u64 s = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    u64 stride = (u64)gridDim.x * blockDim.x;
    for (; s < n; s += stride) {
        u32 v = (u32)(keys[s] & 0xFFFFFFFFULL);
        if (s < n1) { vert[s] = v; slot[v] = (u32)s; hdeg[s] = 0xFFFFFFFFu - (u32)(keys[s] >> 32); } else slot[v] = 0xFFFFFFFFu;
    }
}
__global__ void k_full_order(const u64 *__restrict__ keys, u64 n, u32 *__restrict__ pvert, u32 *__restrict__ pperm) {
    u64 s = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    u64 stride = (u64)gridDim.x * blockDim.x;
    for (; s < n; s += stride) { const u32 v = (u32)(keys[s] & 0xFFFFFFFFULL); pvert[s] = v; pperm[v] = (u32)s; }
}
__global__ void k_relabel_cols(const u32 *__restrict__ j, u64 nnz, const u32 *__restrict__ slot, u32 *__restrict__ jp) {
    u64 q = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    u64 stride = (u64)gridDim.x * blockDim.x;
    for (; q < nnz; q += stride) jp[q] = slot[j[q]];
}
template <int W>
__global__ void k_pack_frontier(const u32 *__restrict__ vert, u64 n1, const u64 *__restrict__ X, u64 *__restrict__ Xp) {
    u64 t = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    u64 stride = (u64)gridDim.x * blockDim.x;
    for (; t < n1 * W; t += stride) {
        u64 s = t / W, w = t % W;
        Xp[t] = X[(u64)vert[s] * W + w];
    }
}

// pack + flops + edges + OR-of-everything in one pass over the frontier (fused_prep): the hot set holds every vertex with
// out-edges, so flops = sum popc(X[v]) * deg(v) and edges = sum [X[v] != 0] * deg(v) over it are the exact totals, and
// G[w] = OR of word column w is the terminal value the pull kernels vote against.  One thread per packed word.
template <int W>
__global__ void __launch_bounds__(256)
k_pack_flops(const u32 *__restrict__ vert, const u32 *__restrict__ hdeg, u64 n1, const u64 *__restrict__ X, u64 *__restrict__ Xp,
             u64 *__restrict__ st /* [0] flops [1] edges */, u64 *__restrict__ G) {
    typedef cub::BlockReduce<u64, 256> Red;
    __shared__ typename Red::TempStorage ts;
    __shared__ u64 sG[W];
    if (threadIdx.x < W) sG[threadIdx.x] = 0;
    __syncthreads();
    const u32 lane = threadIdx.x & 31;
    const u32 gmask = (W >= 32) ? 0xffffffffu : (((1u << W) - 1u) << (lane / W * W));   // the W lanes of one vertex
    u64 fl = 0, ed = 0, g = 0;
    u64 t = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    const u64 stride = (u64)gridDim.x * blockDim.x, total = n1 * W;
    for (u64 base = t - lane; base < total; base += stride) {       // warp-uniform trip count (ballot inside)
        const u64 tt = base + lane;
        u64 x = 0;
        u32 d = 0;
        if (tt < total) {
            const u64 s = tt / W, w = tt % W;
            x = X[(u64)vert[s] * W + w];
            d = hdeg[s];
            Xp[tt] = x;
        }
        const u32 nz = __ballot_sync(0xffffffffu, x != 0) & gmask;
        fl += (u64)__popcll(x) * d;
        if (nz && (lane % W) == 0) ed += d;
        g |= x;
    }
    // OR across the lanes that hold the same word column (lane % W), then one shared atomic per column and warp
#pragma unroll
    for (int o = W; o < 32; o <<= 1) g |= __shfl_xor_sync(0xffffffffu, g, o);
    if (lane < W && g) atomicOr((unsigned long long *)&sG[lane], g);
    u64 tf = Red(ts).Sum(fl);
    __syncthreads();
    u64 te = Red(ts).Sum(ed);
    if (threadIdx.x == 0) {
        if (tf) atomicAdd((unsigned long long *)&st[0], tf);
        if (te) atomicAdd((unsigned long long *)&st[1], te);
    }
    __syncthreads();
    if (threadIdx.x < W && sG[threadIdx.x]) atomicOr((unsigned long long *)&G[threadIdx.x], sG[threadIdx.x]);
}

// the same totals for a frontier that already IS in hot-set order (permuted form): nothing to gather, nothing to write
template <int W>
__global__ void __launch_bounds__(256)
k_ordered_flops(const u32 *__restrict__ hdeg, u64 n1, const u64 *__restrict__ Xp, u64 *__restrict__ st, u64 *__restrict__ G) {
    typedef cub::BlockReduce<u64, 256> Red;
    __shared__ typename Red::TempStorage ts;
    __shared__ u64 sG[W];
    if (threadIdx.x < W) sG[threadIdx.x] = 0;
    __syncthreads();
    const u32 lane = threadIdx.x & 31;
    const u32 gmask = (W >= 32) ? 0xffffffffu : (((1u << W) - 1u) << (lane / W * W));
    u64 fl = 0, ed = 0, g = 0;
    u64 t = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    const u64 stride = (u64)gridDim.x * blockDim.x, total = n1 * W;
    for (u64 base = t - lane; base < total; base += stride) {
        const u64 tt = base + lane;
        u64 x = 0;
        u32 d = 0;
        if (tt < total) { x = Xp[tt]; d = hdeg[tt / W]; }
        const u32 nz = __ballot_sync(0xffffffffu, x != 0) & gmask;
        fl += (u64)__popcll(x) * d;
        if (nz && (lane % W) == 0) ed += d;
        g |= x;
    }
#pragma unroll
    for (int o = W; o < 32; o <<= 1) g |= __shfl_xor_sync(0xffffffffu, g, o);
    if (lane < W && g) atomicOr((unsigned long long *)&sG[lane], g);
    u64 tf = Red(ts).Sum(fl);
    __syncthreads();
    u64 te = Red(ts).Sum(ed);
    if (threadIdx.x == 0) {
        if (tf) atomicAdd((unsigned long long *)&st[0], tf);
        if (te) atomicAdd((unsigned long long *)&st[1], te);
    }
    __syncthreads();
    if (threadIdx.x < W && sG[threadIdx.x]) atomicOr((unsigned long long *)&G[threadIdx.x], sG[threadIdx.x]);
}

void build_hot_pack(const DevCSR &A, const DevCSR &AT, LongRows &lr) {
    u64 n = A.nrows;
    lr.vert.release(); lr.jp.release(); lr.slot.release(); lr.hdeg.release(); lr.n1 = 0; lr.packed = false;
    if (n == 0 || AT.nnz == 0 || n >= 0xFFFFFFFFULL) return;
    DevBuf<u64> keys(n), cnt(1);
    cnt.zero();
    LAUNCH(k_degree_keys, grid_for(n, 256, 148 * 16), 256, 0, A.p.ptr, n, keys.ptr, cnt.ptr);
    sort_keys_u64(keys.ptr, n, 64);
    u64 n1 = read_scalar(cnt.ptr);
    lr.n1 = n1;
    lr.vert.alloc(n1 ? n1 : 1);
    lr.slot.alloc(n);
    lr.hdeg.alloc(n1 ? n1 : 1);
    LAUNCH(k_slots_from_keys, grid_for(n, 256, 148 * 16), 256, 0, keys.ptr, n, n1, lr.vert.ptr, lr.slot.ptr, lr.hdeg.ptr);
    // the order over ALL vertices (the sort already put the sinks last, by id): frontiers may be stored in it (permuted form)
    lr.pvert = std::make_shared<DevBuf<u32>>(n);
    lr.pperm = std::make_shared<DevBuf<u32>>(n);
    LAUNCH(k_full_order, grid_for(n, 256, 148 * 16), 256, 0, keys.ptr, n, lr.pvert->ptr, lr.pperm->ptr);
    static std::atomic<u64> next_tag{1};
    lr.perm_tag = next_tag.fetch_add(1);
    lr.packed = true;
}

// ---- CSR-stream form of A' ----------------------------------------------------------------------------------------
__global__ void k_short_len(const u64 *__restrict__ p, u64 n, u64 LONG, u32 *__restrict__ len_s, u32 *__restrict__ lflag,
                            u64 *__restrict__ maxlong) {
    u64 r = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    u64 stride = (u64)gridDim.x * blockDim.x;
    u64 mx = 0;
    for (; r <= n; r += stride) {
        if (r == n) { len_s[r] = 0; lflag[r] = 0; break; }
        u64 d = p[r + 1] - p[r];
        bool lg = d > LONG;
        len_s[r] = lg ? 0u : (u32)d;
        lflag[r] = lg ? 1u : 0u;
        if (lg && d > mx) mx = d;
    }
    if (mx) atomicMax((unsigned long long *)maxlong, mx);
}
__global__ void k_long_meta(const u64 *__restrict__ p, u64 n, const u32 *__restrict__ lflag, const u64 *__restrict__ lpos,
                            u64 nlong, u32 *__restrict__ lrows, u64 *__restrict__ llen) {
    u64 r = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    u64 stride = (u64)gridDim.x * blockDim.x;
    if (r == 0) llen[nlong] = 0;
    for (; r < n; r += stride)
        if (lflag[r]) { lrows[lpos[r]] = (u32)r; llen[lpos[r]] = p[r + 1] - p[r]; }
}
__global__ void k_fill_stream(const u64 *__restrict__ p, const u32 *__restrict__ j, u64 n, const u32 *__restrict__ slot,
                              const u64 *__restrict__ rp_s, u32 *__restrict__ jp_s, const u32 *__restrict__ lflag,
                              const u64 *__restrict__ lpos, const u64 *__restrict__ lrp, u32 *__restrict__ jp_l) {
    u64 warp = ((u64)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    u64 nwarps = ((u64)gridDim.x * blockDim.x) >> 5;
    u32 lane = threadIdx.x & 31;
    for (u64 r = warp; r < n; r += nwarps) {
        u64 s = p[r], e = p[r + 1];
        u32 *dst = lflag[r] ? (jp_l + lrp[lpos[r]]) : (jp_s + rp_s[r]);
        for (u64 q = lane; q < e - s; q += 32) dst[q] = slot ? slot[j[s + q]] : j[s + q];
    }
}
__global__ void k_window_starts(const u64 *__restrict__ rp_s, u64 n, u64 WIN, u64 nwin, u32 *__restrict__ wstart) {
    u64 k = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (k > nwin) return;
    if (k == nwin) { wstart[k] = (u32)n; return; }
    u64 target = k * WIN, lo = 0, hi = n;      // first row r in [0,n) with rp_s[r] >= target, else n
    while (lo < hi) {
        u64 mid = (lo + hi) >> 1;
        if (rp_s[mid] < target) lo = mid + 1; else hi = mid;
    }
    wstart[k] = (u32)lo;
}

static u64 stream_window(u32 W) { return 1024 / W; }   // window (and short-row limit) in nnz; W <= 4

void build_stream(const DevCSR &AT, LongRows &lr, u32 W) {
    u64 n = AT.nrows;
    u64 WIN = stream_window(W);
    lr.sW = W; lr.swin = WIN; lr.s_packed = lr.packed;
    DevBuf<u32> len_s(n + 1), lflag(n + 1);
    DevBuf<u64> lpos(n + 1), mx(1);
    mx.zero();
    LAUNCH(k_short_len, grid_for(n + 1, 256, 148 * 16), 256, 0, AT.p.ptr, n, WIN, len_s.ptr, lflag.ptr, mx.ptr);
    lr.rp_s.alloc(n + 1);
    exclusive_scan_u32_to_u64(len_s.ptr, lr.rp_s.ptr, n + 1);
    exclusive_scan_u32_to_u64(lflag.ptr, lpos.ptr, n + 1);
    lr.nnz_s = read_scalar(lr.rp_s.ptr + n);
    lr.nlong = read_scalar(lpos.ptr + n);
    lr.maxlong = read_scalar(mx.ptr);
    lr.lrows.alloc(lr.nlong ? lr.nlong : 1);
    lr.lrp.alloc(lr.nlong + 1);
    LAUNCH(k_long_meta, grid_for(n ? n : 1, 256, 148 * 16), 256, 0, AT.p.ptr, n, lflag.ptr, lpos.ptr, lr.nlong, lr.lrows.ptr, lr.lrp.ptr);
    exclusive_scan_u64(lr.lrp.ptr, lr.lrp.ptr, lr.nlong + 1);
    lr.jp_s.alloc(lr.nnz_s ? lr.nnz_s : 1);
    lr.jp_l.alloc(AT.nnz - lr.nnz_s ? AT.nnz - lr.nnz_s : 1);
    if (n) LAUNCH(k_fill_stream, grid_for(n * 32, 256, 148 * 32), 256, 0, AT.p.ptr, AT.j.ptr, n, lr.packed ? lr.slot.ptr : (const u32 *)nullptr,
                  lr.rp_s.ptr, lr.jp_s.ptr, lflag.ptr, lpos.ptr, lr.lrp.ptr, lr.jp_l.ptr);
    lr.nwin = lr.nnz_s / WIN + 1;
    lr.wstart.alloc(lr.nwin + 1);
    LAUNCH(k_window_starts, grid_for(lr.nwin + 1, 256), 256, 0, lr.rp_s.ptr, n, WIN, lr.nwin, lr.wstart.ptr);
}

// flat coalesced gather of one window's nnz into shared memory, then one thread per row reduces its segment
template <int W>
__global__ void __launch_bounds__(128)
k_bits_pull_stream(const u64 *__restrict__ rp_s, const u32 *__restrict__ jp_s, const u32 *__restrict__ wstart,
                   const u64 *__restrict__ Xp, u64 *__restrict__ Y) {
    extern __shared__ u64 sxs[];
    const u32 tid = threadIdx.x;
    u32 r0 = wstart[blockIdx.x], r1 = wstart[blockIdx.x + 1];
    if (r0 == r1) return;
    u64 base = rp_s[r0];
    u32 nnzb = (u32)(rp_s[r1] - base);
    const u32 *cols = jp_s + base;
#pragma unroll 4
    for (u32 e = tid; e < nnzb; e += 128) {
        u32 col = cols[e];
#pragma unroll
        for (int w = 0; w < W; w++) sxs[(size_t)e * W + w] = Xp[(u64)col * W + w];
    }
    __syncthreads();
    for (u32 r = r0 + tid; r < r1; r += 128) {
        u32 s = (u32)(rp_s[r] - base), e = (u32)(rp_s[r + 1] - base);
        u64 acc[W];
#pragma unroll
        for (int w = 0; w < W; w++) acc[w] = 0;
        for (u32 q = s; q < e; q++) {
#pragma unroll
            for (int w = 0; w < W; w++) acc[w] |= sxs[(size_t)q * W + w];
        }
#pragma unroll
        for (int w = 0; w < W; w++) Y[(u64)r * W + w] = acc[w];
    }
}

template <int W>
__global__ void __launch_bounds__(256)
k_bits_pull_stream_long(const u32 *__restrict__ lrows, const u64 *__restrict__ lrp, const u32 *__restrict__ jp_l,
                        const u64 *__restrict__ Xp, u64 *__restrict__ Y) {
    typedef cub::BlockReduce<u64, 256> Red;
    __shared__ typename Red::TempStorage ts;
    u64 s = lrp[blockIdx.x], e = lrp[blockIdx.x + 1];
    u64 c0 = s + (u64)blockIdx.y * LONG_CHUNK;
    if (c0 >= e) return;
    u64 c1 = c0 + LONG_CHUNK;
    if (c1 > e) c1 = e;
    u64 acc[W];
#pragma unroll
    for (int w = 0; w < W; w++) acc[w] = 0;
#pragma unroll 4
    for (u64 q = c0 + threadIdx.x; q < c1; q += 256) {
        u32 k = jp_l[q];
#pragma unroll
        for (int w = 0; w < W; w++) acc[w] |= Xp[(u64)k * W + w];
    }
    u32 row = lrows[blockIdx.x];
#pragma unroll
    for (int w = 0; w < W; w++) {
        u64 r = Red(ts).Reduce(acc[w], OrOp64());
        __syncthreads();
        if (threadIdx.x == 0 && r) atomicOr((unsigned long long *)&Y[(u64)row * W + w], r);
    }
}

void build_long_rows(const DevCSR &AT, LongRows &lr) {
    u64 n = AT.nrows;
    lr.rows.release();
    lr.n = 0;
    lr.built = true;
    if (n == 0) return;
    DevBuf<u32> flag(n + 1);
    DevBuf<u64> pos(n + 1), mx(1);
    mx.zero();
    LAUNCH(k_flag_long, grid_for(n, 256, 1 << 16), 256, 0, AT.p.ptr, n, flag.ptr, mx.ptr);
    CUDA_TRY(cudaMemsetAsync(flag.ptr + n, 0, sizeof(u32), stream()));
    exclusive_scan_u32_to_u64(flag.ptr, pos.ptr, n + 1);
    u64 nl = read_scalar(pos.ptr + n);
    lr.n = nl;
    lr.maxdeg = read_scalar(mx.ptr);
    if (nl) {
        lr.rows.alloc(nl);
        LAUNCH(k_scatter_flagged, grid_for(n, 256, 1 << 16), 256, 0, flag.ptr, pos.ptr, n, lr.rows.ptr);
        lr.choff.alloc(nl + 1);
        LAUNCH(k_long_chunks, grid_for(nl + 1, 256), 256, 0, lr.rows.ptr, nl, AT.p.ptr, lr.choff.ptr);
        exclusive_scan_u64(lr.choff.ptr, lr.choff.ptr, nl + 1);
        lr.nchunks = read_scalar(lr.choff.ptr + nl);
    }
    build_mid_list(AT, lr);
    if (lr.packed) {
        lr.jp.alloc(AT.nnz);
        LAUNCH(k_relabel_cols, grid_for(AT.nnz, 256, 148 * 32), 256, 0, AT.j.ptr, AT.nnz, lr.slot.ptr, lr.jp.ptr);
    }
    u64 ndiag = (n + AT.nnz) / 256 + 2;
    lr.mp_r.alloc(ndiag);
    LAUNCH(k_mp_coords, grid_for(ndiag, 256), 256, 0, AT.p.ptr, n, AT.nnz, ndiag, lr.mp_r.ptr);
}

// persisting-L2 access-policy window on the library stream (cudaLimitPersistingL2CacheSize is set at context bring-up)
static u64 g_l2_carved = 0;        // bytes of L2 currently set aside for persisting accesses
static void carve_l2(u64 bytes) {
    if (bytes == g_l2_carved) return;
    CUDA_TRY(cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, bytes));
    g_l2_carved = bytes;
}
static void set_l2_window(const void *base, u64 bytes) {
    Context &cx = ctx();
    if (cx.l2_persist_max == 0) return;
    carve_l2(std::min<u64>(bytes, cx.l2_persist_max));
    cudaStreamAttrValue a;
    memset(&a, 0, sizeof(a));
    u64 nb = bytes;
    if (nb > cx.l2_window_max) nb = cx.l2_window_max;
    a.accessPolicyWindow.base_ptr = const_cast<void *>(base);
    a.accessPolicyWindow.num_bytes = nb;
    a.accessPolicyWindow.hitRatio = nb <= g_l2_carved ? 1.0f : (float)((double)g_l2_carved / (double)nb);
    a.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
    a.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
    CUDA_TRY(cudaStreamSetAttribute(stream(), cudaStreamAttributeAccessPolicyWindow, &a));
}
static void clear_l2_window() {
    Context &cx = ctx();
    if (cx.l2_persist_max == 0) return;
    cudaStreamAttrValue a;
    memset(&a, 0, sizeof(a));
    a.accessPolicyWindow.num_bytes = 0;
    a.accessPolicyWindow.hitRatio = 0.0f;
    a.accessPolicyWindow.hitProp = cudaAccessPropertyNormal;
    a.accessPolicyWindow.missProp = cudaAccessPropertyNormal;
    CUDA_TRY(cudaStreamSetAttribute(stream(), cudaStreamAttributeAccessPolicyWindow, &a));
    if (cx.opt_l2_reset) CUDA_TRY(cudaCtxResetPersistingL2Cache());
    if (cx.opt_l2_reset >= 2) carve_l2(0);       // hand the set-aside back to the other kernels until the next pull
}

template <int W>
static void hop_impl(const DevBits &Xin, const DevCSR &A, const DevCSR *AT, LongRows *lr, DevBits &Y,
                     u64 *flops_out, int *path_out) {
    Context &cx = ctx();
    u64 n = A.nrows, m = A.ncols;
    // X in permuted form: usable as it stands only by the pull through the tables it was ordered for; anything else works on a
    // natural-order copy (cannot happen on the bench path: the caller orders a frontier only for the operand it is pushed through)
    DevBits Xnat;
    const bool ordered = Xin.permuted() && AT && lr && lr->packed && Xin.perm_tag == lr->perm_tag && lr->jp.ptr && lr->hdeg.ptr &&
                         cx.opt_pull_mode != 0 && cx.opt_hot_pack && cx.opt_pull_kernel >= 4 && n == m;
    const DevBits *Xuse = &Xin;
    if (Xin.permuted() && !ordered) { bits_copy(Xin, Xnat); bits_naturalise(Xnat); Xuse = &Xnat; }
    const DevBits &X = *Xuse;
    Y.clear();
    Y.nrows = X.nrows; Y.ncols = m; Y.W = W;
    Y.w.alloc(m * W);
    DevBuf<u32> flag(n + 1);
    DevBuf<u64> st(2);
    st.zero();
    // fused_prep: when the hot-set tables exist, one pass packs the frontier for the pull AND yields flops / edges / the OR
    // terminal; the separate flops pass (needed for its per-vertex flags) then only runs if the push direction wins
    const bool fused = cx.opt_fused_prep && AT && lr && cx.opt_hot_pack && lr->packed && lr->jp.ptr && lr->hdeg.ptr &&
                       cx.opt_pull_mode != 0 && cx.opt_pull_kernel != 2 && n == m;
    DevBuf<u64> Xp, Gd;
    if (ordered) {            // already in gather order: totals only
        Gd.alloc(W);
        Gd.zero();
        if (lr->n1) LAUNCH((k_ordered_flops<W>), grid_for(lr->n1 * W, 256, 148 * 16), 256, 0, lr->hdeg.ptr, lr->n1, X.w.ptr, st.ptr, Gd.ptr);
    } else if (fused) {
        Xp.alloc(lr->n1 * W);
        Gd.alloc(W);
        Gd.zero();
        if (lr->n1) LAUNCH((k_pack_flops<W>), grid_for(lr->n1 * W, 256, 148 * 16), 256, 0, lr->vert.ptr, lr->hdeg.ptr, lr->n1, X.w.ptr, Xp.ptr, st.ptr, Gd.ptr);
    } else if (n) LAUNCH(k_bits_flops, grid_for(n, 256, 148 * 16), 256, 0, X.w.ptr, (u32)W, n, A.p.ptr, flag.ptr, st.ptr);
    u64 hst[2] = {0, 0};
    d2h(hst, st.ptr, 2);
    sync_stream();
    if (flops_out) *flops_out = hst[0];
    u64 edges = hst[1];
    bool pull = false;
    if (AT && lr) {
        if (cx.opt_pull_mode == 1) pull = true;
        else if (cx.opt_pull_mode == 0) pull = false;
        else pull = edges * 4 > A.nnz; // direction switch: frontier touches > 1/4 of the edges
    }
    if (edges == 0) { Y.w.zero(); if (path_out) *path_out = 0; return; }
    const u32 *gj = AT ? AT->j.ptr : nullptr;   // gather index stream
    const u64 *gx = X.w.ptr;                    // gather source
    u64 gn = n;                                 // gather footprint in vertices
    if (ordered && !pull) {                     // the push direction works on natural order: undo the frontier's order and start over
        DevBits Xn;
        bits_copy(Xin, Xn);
        bits_naturalise(Xn);
        hop_impl<W>(Xn, A, AT, lr, Y, flops_out, path_out);
        return;
    }
    if (fused && !ordered && !pull) {           // push needs the per-vertex active flags of the classic pass
        DevBuf<u64> st2(2);
        st2.zero();
        LAUNCH(k_bits_flops, grid_for(n, 256, 148 * 16), 256, 0, X.w.ptr, (u32)W, n, A.p.ptr, flag.ptr, st2.ptr);
    }
    const bool stream_kernel = pull && W <= 4 && cx.opt_pull_kernel == 2;
    if (pull) {   // auxiliary tables are built once per (immutable) matrix, lazily
        if (!lr->packed && cx.opt_hot_pack) build_hot_pack(A, *AT, *lr);
        if (stream_kernel) { if (lr->sW != (u32)W || lr->s_packed != lr->packed) build_stream(*AT, *lr, W); }
        else if (!lr->built || (lr->packed && !lr->jp.ptr)) build_long_rows(*AT, *lr);
    }
    if (stream_kernel) {
        if (lr->s_packed) {
            Xp.alloc(lr->n1 * W);
            if (lr->n1) LAUNCH((k_pack_frontier<W>), grid_for(lr->n1 * W, 256, 148 * 32), 256, 0, lr->vert.ptr, lr->n1, X.w.ptr, Xp.ptr);
            gx = Xp.ptr; gn = lr->n1;
        }
        const size_t smem = (size_t)2 * lr->swin * W * sizeof(u64);
        {
            TimedScope ts(TK_BITS_PULL, 4 * AT->nnz + 8 * (m + 1) + 8ULL * W * gn + 8ULL * W * m);
            LAUNCH((k_bits_pull_stream<W>), (u32)lr->nwin, 128, smem, lr->rp_s.ptr, lr->jp_s.ptr, lr->wstart.ptr, gx, Y.w.ptr);
        }
        if (lr->nlong) {
            dim3 g((u32)lr->nlong, (u32)((lr->maxlong + LONG_CHUNK - 1) / LONG_CHUNK));
            TimedScope ts(TK_BITS_PULL_LONG, 0);
            LAUNCH((k_bits_pull_stream_long<W>), g, 256, 0, lr->lrows.ptr, lr->lrp.ptr, lr->jp_l.ptr, gx, Y.w.ptr);
        }
        if (path_out) *path_out = 5;
        return;
    }
    if (ordered) {                              // the frontier already is the packed gather source (its first n1 records)
        gj = lr->jp.ptr; gx = X.w.ptr; gn = lr->n1;
    } else if (pull && cx.opt_hot_pack && lr->packed && lr->jp.ptr) {
        if (!fused) {
            Xp.alloc(lr->n1 * W);
            if (lr->n1) LAUNCH((k_pack_frontier<W>), grid_for(lr->n1 * W, 256, 148 * 32), 256, 0, lr->vert.ptr, lr->n1, X.w.ptr, Xp.ptr);
        }
        gj = lr->jp.ptr; gx = Xp.ptr; gn = lr->n1;
    }
    const u64 *Gp = nullptr;
    // opt_early_exit: 0 never, 1 auto, 2 always.  Measured on B200 (RMAT-24, hop 3): the in-loop vote pays for itself at
    // W = 2 and 4 (hub rows saturate: pull 2.92 -> 2.56 ms and 4.07 -> 3.66 ms), costs more than it saves at W = 1
    // (2.08 -> 3.14 ms, few rows ever hold all 64 bits) and at W >= 8; very dense frontiers always benefit.
    const bool dense_frontier = hst[0] >= (hst[1] * X.nrows) / 4;   // flops/edges = degree-weighted mean popcount of a gathered word
    if (pull && (cx.opt_early_exit == 2 || (cx.opt_early_exit == 1 && (dense_frontier || W == 2 || W == 4)))) {
        if (!((fused && gx == Xp.ptr) || ordered)) {   // the fused / ordered pass already left the OR of the packed frontier in Gd
            Gd.alloc(W);
            Gd.zero();
            if (gn) LAUNCH((k_or_all<W>), grid_for(gn, 256, 148 * 8), 256, 0, gx, gn, Gd.ptr);
        }
        Gp = Gd.ptr;
    }
    if (pull && cx.opt_pull_kernel == 1) {
        constexpr int TILE = MpCfg<W>::TILE;
        const size_t smem = ((size_t)TILE * W + (size_t)(TILE + 1) * W) * sizeof(u64) + (size_t)(TILE + 2) * sizeof(u32);
        static bool attr_set = false;
        if (!attr_set) {
            CUDA_TRY(cudaFuncSetAttribute(k_bits_pull_mp<W>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            attr_set = true;
        }
        Y.w.zero();
        u64 total = m + AT->nnz;
        u64 ntiles = (total + TILE - 1) / TILE;
        TimedScope ts(TK_BITS_PULL, 4 * AT->nnz + 8 * (m + 1) + 8ULL * W * gn + 8ULL * W * m);
        LAUNCH((k_bits_pull_mp<W>), (u32)ntiles, 256, smem, AT->p.ptr, gj, m, AT->nnz, gx, Y.w.ptr, lr->mp_r.ptr);
        if (path_out) *path_out = 4;
    } else if (pull) {
        {
            // compulsory traffic: stream A' col_idx + rowptr, read X once, write Y once (X gathers hit L2)
            TimedScope ts(TK_BITS_PULL, 4 * AT->nnz + 8 * (m + 1) + 8ULL * W * gn + 8ULL * W * m);
            // early termination only pays when the gathered words are dense (degree-weighted mean popcount >= 1/4 of the rows)
            const bool early = Gp != nullptr;
            u64 totb = gn * W * 8;
            const u32 tot_bytes = totb > 0xFFFFFFF0ULL ? 0xFFFFFFF0u : (u32)totb;
            const u32 hot_bytes = (u64)cx.opt_hot_bytes < tot_bytes ? (u32)cx.opt_hot_bytes : tot_bytes;
            // grid = every CTA resident at once (occupancy x SMs) unless pull_grid overrides it: rows are dealt round-robin to
            // 8-lane groups, so one full wave keeps all SMs busy to the end (measured: 3.41 ms at 16 CTAs/SM, 3.12 ms resident)
            if (cx.opt_pull_kernel >= 5) {
                if (!lr->seg_built) build_seg_list(*AT, *lr);
                auto occ = [&](auto kern) {
                    int per_sm = 0;
                    CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 256, 0));
                    if (per_sm < 1) per_sm = 1;
                    if (cx.opt_pull_grid > 0) per_sm = (int)cx.opt_pull_grid;
                    return (u32)cx.num_sms * (u32)per_sm;
                };
                // keep the hot prefix of the packed frontier L2-resident for the duration of the pull: a persisting access-policy
                // window on the library stream (no per-instruction hint exists for 256-bit loads)
                const bool window = cx.opt_l2_window > 0 && (gx == Xp.ptr || ordered) && gn;
                if (window) set_l2_window(gx, std::min<u64>((u64)cx.opt_l2_window, gn * W * 8));
                auto small = [&](auto kern, u64 split) {
                    LAUNCH(kern, grid_for(m * split, 256, (u64)cx.num_sms * 32), 256, 0, AT->p.ptr, gj, m, gx, Y.w.ptr, hot_bytes, tot_bytes);
                };
                constexpr int SP = PullCfg<W>::SPLIT;
                auto seg = [&](auto kern) {
                    if (lr->ns) LAUNCH(kern, occ(kern), 256, 0, lr->s_row.ptr, lr->s_start.ptr, lr->s_len.ptr, lr->ns, gj, gx, Y.w.ptr, Gp, hot_bytes, tot_bytes);
                };
#define SEG_LAUNCH(H) do { \
        if (cx.opt_unroll >= 8) { if (early) seg(k_pull_seg<W, H, 8, true>); else seg(k_pull_seg<W, H, 8, false>); } \
        else if (cx.opt_unroll >= 4) { if (early) seg(k_pull_seg<W, H, 4, true>); else seg(k_pull_seg<W, H, 4, false>); } \
        else { if (early) seg(k_pull_seg<W, H, 2, true>); else seg(k_pull_seg<W, H, 2, false>); } } while (0)
#define SMALL_LAUNCH(H) do { if (cx.opt_small_split && SP > 1) small(k_pull_small<W, H, SP>, SP); else small(k_pull_small<W, H, 1>, 1); } while (0)
                if (cx.opt_hints > 0 || (cx.opt_hints < 0 && W <= 2)) { SMALL_LAUNCH(true); SEG_LAUNCH(true); }
                else { SMALL_LAUNCH(false); SEG_LAUNCH(false); }
#undef SMALL_LAUNCH
#undef SEG_LAUNCH
                if (window) clear_l2_window();
                if (path_out) *path_out = 3;
                return;
            } else if (cx.opt_pull_kernel == 4) {
                auto occ = [&](auto kern) {
                    int per_sm = 0;
                    CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 256, 0));
                    if (per_sm < 1) per_sm = 1;
                    if (cx.opt_pull_grid > 0) per_sm = (int)cx.opt_pull_grid;
                    return (u32)cx.num_sms * (u32)per_sm;
                };
                auto small = [&](auto kern) { LAUNCH(kern, grid_for(m, 256, (u64)cx.num_sms * 32), 256, 0, AT->p.ptr, gj, m, gx, Y.w.ptr, hot_bytes, tot_bytes); };
                auto mid = [&](auto kern) {
                    if (lr->nm) LAUNCH(kern, occ(kern), 256, 0, lr->m_row.ptr, lr->m_start.ptr, lr->m_len.ptr, lr->nm, gj, gx, Y.w.ptr, Gp, hot_bytes, tot_bytes);
                };
#define MID_LAUNCH(H) do { \
        if (cx.opt_unroll >= 8) { if (early) mid(k_bits_pull_mid<W, H, 8, true>); else mid(k_bits_pull_mid<W, H, 8, false>); } \
        else if (cx.opt_unroll >= 4) { if (early) mid(k_bits_pull_mid<W, H, 4, true>); else mid(k_bits_pull_mid<W, H, 4, false>); } \
        else { if (early) mid(k_bits_pull_mid<W, H, 2, true>); else mid(k_bits_pull_mid<W, H, 2, false>); } } while (0)
                if (cx.opt_hints > 0 || (cx.opt_hints < 0 && W <= 2)) { small(k_bits_pull_small<W, true>); MID_LAUNCH(true); }
                else { small(k_bits_pull_small<W, false>); MID_LAUNCH(false); }
#undef MID_LAUNCH
            } else {
            auto go = [&](auto kern) {
                int per_sm = 0;
                CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 256, 0));
                if (per_sm < 1) per_sm = 1;
                if (cx.opt_pull_grid > 0) per_sm = (int)cx.opt_pull_grid;
                const u32 g = (u32)cx.num_sms * (u32)per_sm;
                LAUNCH(kern, g, 256, 0, AT->p.ptr, gj, m, gx, Y.w.ptr, Gp, hot_bytes, tot_bytes);
            };
#define PULL_LAUNCH(H, UU) do { \
        if (cx.opt_pull_kernel == 0) { if (early) go(k_bits_pull<W, H, UU, true>); else go(k_bits_pull<W, H, UU, false>); } \
        else { if (early) go(k_bits_pull_pipe<W, H, UU, true>); else go(k_bits_pull_pipe<W, H, UU, false>); } } while (0)
            if (cx.opt_hints > 0 || (cx.opt_hints < 0 && W <= 2)) {
                if (cx.opt_unroll >= 4) PULL_LAUNCH(true, 4); else if (cx.opt_unroll >= 2) PULL_LAUNCH(true, 2); else PULL_LAUNCH(true, 1);
            } else {
                if (cx.opt_unroll >= 4) PULL_LAUNCH(false, 4); else if (cx.opt_unroll >= 2) PULL_LAUNCH(false, 2); else PULL_LAUNCH(false, 1);
            }
#undef PULL_LAUNCH
            }
        }
        if (lr->n) {
            TimedScope ts(TK_BITS_PULL_LONG, 0);
            LAUNCH((k_bits_pull_long<W>), (u32)lr->nchunks, 256, 0, lr->rows.ptr, lr->choff.ptr, (u32)lr->n, AT->p.ptr, gj, gx, Y.w.ptr, Gp);
        }
        if (path_out) *path_out = 3;
    } else {
        Y.w.zero();
        CUDA_TRY(cudaMemsetAsync(flag.ptr + n, 0, sizeof(u32), stream()));
        DevBuf<u64> pos(n + 1);
        exclusive_scan_u32_to_u64(flag.ptr, pos.ptr, n + 1);
        u64 nact = read_scalar(pos.ptr + n);
        DevBuf<u32> act(nact);
        DevBuf<u64> cum(nact + 1), astart(nact);
        LAUNCH(k_compact_active, grid_for(n, 256, 1 << 16), 256, 0, flag.ptr, pos.ptr, n, A.p.ptr, act.ptr, cum.ptr, astart.ptr);
        CUDA_TRY(cudaMemsetAsync(cum.ptr + nact, 0, sizeof(u64), stream()));
        exclusive_scan_u64(cum.ptr, cum.ptr, nact + 1);
        u64 nchunks = (edges + PUSH_CHUNK - 1) / PUSH_CHUNK;
        {
            // col_idx segments + per-active (id,cum,start) + X words + one RED per (edge, word)
            TimedScope ts(TK_BITS_PUSH, 4 * edges + 20 * nact + 8ULL * W * nact + 8ULL * W * edges);
            LAUNCH((k_bits_push<W>), (u32)nchunks, 256, 0, act.ptr, cum.ptr, astart.ptr, nact, edges, A.j.ptr, X.w.ptr, Y.w.ptr);
        }
        if (path_out) *path_out = 2;
    }
}

// builds every per-matrix table of the pull direction now (B200_Matrix_prepare) instead of inside the first hop that pulls
void bits_prepare_pull(const DevCSR &A, const DevCSR &AT, LongRows &lr) {
    if (A.nrows != A.ncols || !AT.nnz) return;
    Context &cx = ctx();
    if (!lr.packed && cx.opt_hot_pack) build_hot_pack(A, AT, lr);
    if (!lr.built || (lr.packed && !lr.jp.ptr)) build_long_rows(AT, lr);
    if (cx.opt_pull_kernel >= 5 && !lr.seg_built) build_seg_list(AT, lr);
}

void bits_hop(const DevBits &X, const DevCSR &A, const DevCSR *AT, LongRows *lr, DevBits &Y, u64 *flops_out,
              int *path_out) {
    if (X.ncols != A.nrows) throw GrbError(-6, "mxm: inner dimensions differ");
    switch (X.W) {
    case 1: hop_impl<1>(X, A, AT, lr, Y, flops_out, path_out); break;
    case 2: hop_impl<2>(X, A, AT, lr, Y, flops_out, path_out); break;
    case 4: hop_impl<4>(X, A, AT, lr, Y, flops_out, path_out); break;
    case 8: hop_impl<8>(X, A, AT, lr, Y, flops_out, path_out); break;
    case 16: hop_impl<16>(X, A, AT, lr, Y, flops_out, path_out); break;
    default: throw GrbError(-8, "bit-frontier: unsupported word count");
    }
}

// ---------------------------------------------------------------------------- elementwise
__global__ void k_bits_andnot(u64 *__restrict__ y, const u64 *__restrict__ m, u64 n) {
    u64 t = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    u64 stride = (u64)gridDim.x * blockDim.x;
    for (; t < n; t += stride) y[t] &= ~m[t];
}
__global__ void k_bits_or(u64 *__restrict__ y, const u64 *__restrict__ z, u64 n) {
    u64 t = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    u64 stride = (u64)gridDim.x * blockDim.x;
    for (; t < n; t += stride) y[t] |= z[t];
}
void bits_andnot(DevBits &Y, const DevBits &M) {
    require_natural(Y, "bits_andnot"); require_natural(M, "bits_andnot");
    if (Y.W != M.W || Y.ncols != M.ncols) throw GrbError(-6, "bits_andnot: shape mismatch");
    u64 n = Y.ncols * Y.W;
    if (n) LAUNCH(k_bits_andnot, grid_for(n, 256, 148 * 16), 256, 0, Y.w.ptr, M.w.ptr, n);
}
void bits_or(DevBits &Y, const DevBits &Z) {
    require_natural(Y, "bits_or"); require_natural(Z, "bits_or");
    if (Y.W != Z.W || Y.ncols != Z.ncols) throw GrbError(-6, "bits_or: shape mismatch");
    u64 n = Y.ncols * Y.W;
    if (n) LAUNCH(k_bits_or, grid_for(n, 256, 148 * 16), 256, 0, Y.w.ptr, Z.w.ptr, n);
}
void bits_copy(const DevBits &X, DevBits &Y) {
    Y.clear();
    Y.nrows = X.nrows; Y.ncols = X.ncols; Y.W = X.W;
    Y.w.alloc(X.ncols * X.W);
    d2d(Y.w.ptr, X.w.ptr, X.ncols * X.W);
    Y.pvert = X.pvert; Y.perm_tag = X.perm_tag;       // a copy keeps the vertex order of its source
}

} // namespace b200
