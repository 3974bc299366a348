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

namespace b200 {

static const u64 PUSH_CHUNK = 8192;
static const u64 LONG_ROW = 4096;   // pull: rows longer than this are split over several CTAs
static const u64 LONG_CHUNK = 8192;

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
// One CTA = 1024 consecutive vertices (32 warps).  For every frontier row (bit) the CTA counts /
// ranks its set vertices with warp ballots; counts are laid out row-major-by-tile so ONE global
// exclusive scan yields final CSR positions, ascending in vertex id within each row.
template <bool FILL>
__global__ void __launch_bounds__(1024)
k_bits_tiles(const u64 *__restrict__ X, u64 n, u32 W, u64 ntiles, u32 *__restrict__ tc,
             const u64 *__restrict__ off, u32 *__restrict__ Cj) {
    __shared__ unsigned short wc[32][64];
    const u32 tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    u64 tile = blockIdx.x;
    u64 v = tile * 1024 + tid;
    for (u32 w = 0; w < W; w++) {
        u64 word = (v < n) ? X[v * W + w] : 0ULL;
        int any = __syncthreads_or(word != 0ULL);
        if (!any) {
            if (!FILL && tid < 64) tc[((u64)w * 64 + tid) * ntiles + tile] = 0;
            continue;
        }
        u32 c0 = 0, c1 = 0;
        u32 wany = __ballot_sync(0xffffffffu, word != 0ULL);
        if (wany) {
#pragma unroll 8
            for (u32 b = 0; b < 64; b++) {
                u32 m = __ballot_sync(0xffffffffu, (word >> b) & 1ULL);
                if ((b & 31) == lane) { if (b < 32) c0 = __popc(m); else c1 = __popc(m); }
            }
        }
        wc[warp][lane] = (unsigned short)c0;
        wc[warp][lane + 32] = (unsigned short)c1;
        __syncthreads();
        if (tid < 64) {
            u32 run = 0;
            for (u32 q = 0; q < 32; q++) { u32 t = wc[q][tid]; wc[q][tid] = (unsigned short)run; run += t; }
            if (!FILL) tc[((u64)w * 64 + tid) * ntiles + tile] = run;
        }
        __syncthreads();
        if (FILL && wany) {
            u32 lt = (1u << lane) - 1u;
            for (u32 b = 0; b < 64; b++) {
                u32 bit = (u32)((word >> b) & 1ULL);
                u32 m = __ballot_sync(0xffffffffu, bit);
                if (bit) {
                    u64 o = off[((u64)w * 64 + b) * ntiles + tile] + wc[warp][b] + __popc(m & lt);
                    Cj[o] = (u32)v;
                }
            }
        }
        __syncthreads();
    }
}

__global__ void k_bits_rowptr(const u64 *__restrict__ off, u64 nrows, u64 ntiles, u64 *__restrict__ Cp) {
    u64 r = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (r <= nrows) Cp[r] = off[r * ntiles];
}

void bits_to_csr(const DevBits &X, DevCSR &C) {
    u64 n = X.ncols;
    u32 W = X.W;
    C.clear();
    C.nrows = X.nrows; C.ncols = n;
    C.p.alloc(X.nrows + 1);
    if (n == 0) { C.p.zero(); C.nnz = 0; return; }
    u64 ntiles = (n + 1023) / 1024;
    u64 ncnt = (u64)64 * W * ntiles;
    DevBuf<u32> tc(ncnt + 1);
    DevBuf<u64> off(ncnt + 1);
    {
        TimedScope ts(TK_BITS_COUNT, 8ULL * W * n);
        LAUNCH((k_bits_tiles<false>), (u32)ntiles, 1024, 0, X.w.ptr, n, W, ntiles, tc.ptr, (const u64 *)nullptr, (u32 *)nullptr);
    }
    CUDA_TRY(cudaMemsetAsync(tc.ptr + ncnt, 0, sizeof(u32), stream()));
    exclusive_scan_u32_to_u64(tc.ptr, off.ptr, ncnt + 1);
    LAUNCH(k_bits_rowptr, grid_for(X.nrows + 1, 256), 256, 0, off.ptr, X.nrows, ntiles, C.p.ptr);
    u64 nnz = read_scalar(off.ptr + X.nrows * ntiles);
    C.nnz = nnz;
    C.j.alloc(nnz);
    if (nnz) {
        TimedScope ts(TK_BITS_FILL, 8ULL * W * n + 4 * nnz);
        LAUNCH((k_bits_tiles<true>), (u32)ntiles, 1024, 0, X.w.ptr, n, W, ntiles, (u32 *)nullptr, off.ptr, C.j.ptr);
    }
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

// ---------------------------------------------------------------------------- pull
// 8 lanes per output vertex j: OR of X[k] over k in A'(j,:).  Rows longer than LONG_ROW are
// zeroed here and finished by k_bits_pull_long (several CTAs per row, RED.OR into Y).
template <int W>
__global__ void __launch_bounds__(256)
k_bits_pull(const u64 *__restrict__ ATp, const u32 *__restrict__ ATj, u64 n, const u64 *__restrict__ X,
            u64 *__restrict__ Y) {
    const u32 lane8 = threadIdx.x & 7;
    u64 group = ((u64)blockIdx.x * blockDim.x + threadIdx.x) >> 3;
    u64 ngroups = ((u64)gridDim.x * blockDim.x) >> 3;
    u64 warp_first = group - ((threadIdx.x & 31) >> 3); // group id of this warp's first 8-lane group
    for (u64 base = warp_first; base < n; base += ngroups) {
        u64 j = base + ((threadIdx.x & 31) >> 3);
        u64 s = 0, e = 0;
        if (j < n) { s = ATp[j]; e = ATp[j + 1]; }
        if (e - s > LONG_ROW) e = s;
        u64 acc[W];
#pragma unroll
        for (int w = 0; w < W; w++) acc[w] = 0;
        for (u64 q = s + lane8; q < e; q += 8) {
            u32 k = ATj[q];
#pragma unroll
            for (int w = 0; w < W; w++) acc[w] |= X[(u64)k * W + w];
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
                if ((w & 7) == (int)lane8) Y[j * W + w] = acc[w];
        }
    }
}

struct OrOp64 { __device__ u64 operator()(u64 a, u64 b) const { return a | b; } };

template <int W>
__global__ void __launch_bounds__(256)
k_bits_pull_long(const u32 *__restrict__ lrows, const u64 *__restrict__ ATp, const u32 *__restrict__ ATj,
                 const u64 *__restrict__ X, u64 *__restrict__ Y) {
    typedef cub::BlockReduce<u64, 256> Red;
    __shared__ typename Red::TempStorage ts;
    u32 j = lrows[blockIdx.x];
    u64 s = ATp[j], e = ATp[j + 1];
    u64 c0 = s + (u64)blockIdx.y * LONG_CHUNK;
    if (c0 >= e) return;
    u64 c1 = c0 + LONG_CHUNK;
    if (c1 > e) c1 = e;
    u64 acc[W];
#pragma unroll
    for (int w = 0; w < W; w++) acc[w] = 0;
    for (u64 q = c0 + threadIdx.x; q < c1; q += 256) {
        u32 k = ATj[q];
#pragma unroll
        for (int w = 0; w < W; w++) acc[w] |= X[(u64)k * W + w];
    }
#pragma unroll
    for (int w = 0; w < W; w++) {
        u64 r = Red(ts).Reduce(acc[w], OrOp64());
        __syncthreads();
        if (threadIdx.x == 0 && r) atomicOr((unsigned long long *)&Y[(u64)j * W + w], r);
    }
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
__global__ void k_scatter_flagged(const u32 *__restrict__ flag, const u64 *__restrict__ pos, u64 n, u32 *__restrict__ out) {
    u64 r = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    u64 stride = (u64)gridDim.x * blockDim.x;
    for (; r < n; r += stride)
        if (flag[r]) out[pos[r]] = (u32)r;
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
    }
}

template <int W>
static void hop_impl(const DevBits &X, const DevCSR &A, const DevCSR *AT, const LongRows *lr, DevBits &Y,
                     u64 *flops_out, int *path_out) {
    Context &cx = ctx();
    u64 n = A.nrows, m = A.ncols;
    Y.clear();
    Y.nrows = X.nrows; Y.ncols = m; Y.W = W;
    Y.w.alloc(m * W);
    DevBuf<u32> flag(n + 1);
    DevBuf<u64> st(2);
    st.zero();
    if (n) LAUNCH(k_bits_flops, grid_for(n, 256, 148 * 16), 256, 0, X.w.ptr, (u32)W, n, A.p.ptr, flag.ptr, st.ptr);
    u64 hst[2] = {0, 0};
    d2h(hst, st.ptr, 2);
    sync_stream();
    if (flops_out) *flops_out = hst[0];
    u64 edges = hst[1];
    bool pull = false;
    if (AT && lr && lr->built) {
        if (cx.opt_pull_mode == 1) pull = true;
        else if (cx.opt_pull_mode == 0) pull = false;
        else pull = edges * 4 > A.nnz; // direction switch: frontier touches > 1/4 of the edges
    }
    if (edges == 0) { Y.w.zero(); if (path_out) *path_out = 0; return; }
    if (pull) {
        u32 grid = (u32)cx.num_sms * 16;
        {
            // compulsory traffic: stream A' col_idx + rowptr, read X once, write Y once (X gathers hit L2)
            TimedScope ts(TK_BITS_PULL, 4 * AT->nnz + 8 * (m + 1) + 8ULL * W * n + 8ULL * W * m);
            LAUNCH((k_bits_pull<W>), grid, 256, 0, AT->p.ptr, AT->j.ptr, m, X.w.ptr, Y.w.ptr);
        }
        if (lr->n) {
            u32 gy = (u32)((lr->maxdeg + LONG_CHUNK - 1) / LONG_CHUNK);
            dim3 g((u32)lr->n, gy);
            TimedScope ts(TK_BITS_PULL_LONG, 0);
            LAUNCH((k_bits_pull_long<W>), g, 256, 0, lr->rows.ptr, AT->p.ptr, AT->j.ptr, X.w.ptr, Y.w.ptr);
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

void bits_hop(const DevBits &X, const DevCSR &A, const DevCSR *AT, const LongRows *lr, DevBits &Y, u64 *flops_out,
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
    if (Y.W != M.W || Y.ncols != M.ncols) throw GrbError(-6, "bits_andnot: shape mismatch");
    u64 n = Y.ncols * Y.W;
    if (n) LAUNCH(k_bits_andnot, grid_for(n, 256, 148 * 16), 256, 0, Y.w.ptr, M.w.ptr, n);
}
void bits_or(DevBits &Y, const DevBits &Z) {
    if (Y.W != Z.W || Y.ncols != Z.ncols) throw GrbError(-6, "bits_or: shape mismatch");
    u64 n = Y.ncols * Y.W;
    if (n) LAUNCH(k_bits_or, grid_for(n, 256, 148 * 16), 256, 0, Y.w.ptr, Z.w.ptr, n);
}
void bits_copy(const DevBits &X, DevBits &Y) {
    Y.clear();
    Y.nrows = X.nrows; Y.ncols = X.ncols; Y.W = X.W;
    Y.w.alloc(X.ncols * X.W);
    d2d(Y.w.ptr, X.w.ptr, X.ncols * X.W);
}

} // namespace b200
