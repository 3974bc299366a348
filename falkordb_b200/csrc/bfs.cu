// bfs.cu -- single-source BFS level / parent, the kernel behind LAGr_BreadthFirstSearch_Extended as the
// reference consumes it (graph/src/runtime/functions/algo_procedures.rs:1079-1148): level(src)=0,
// parent(src)=src, unreached = -1, optional max_level cap.  LAGraph's parent is "any" valid parent
// (ANY_SECONDI); this implementation makes it deterministic: the MINIMUM parent id in the previous
// level (RED.MIN on a 64-bit candidate array), same rule as oracle/grb_oracle.c orc_bfs.
// Frontier expansion is the load-balanced flat expansion used by the mxm push kernels.
// Algorithmic bytes: 4*m_visited*2 + 16*n_visited + 16*n_reached.
#include "common.cuh"
#include "ops.cuh"

namespace b200 {

static const u64 BFS_CHUNK = 8192;

__global__ void k_bfs_init(i64 *__restrict__ level, i64 *__restrict__ parent, u64 *__restrict__ cand, u64 n, u64 src) {
    u64 t = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    u64 stride = (u64)gridDim.x * blockDim.x;
    for (; t < n; t += stride) {
        level[t] = (t == src) ? 0 : -1;
        if (parent) parent[t] = (t == src) ? (i64)src : -1;
        cand[t] = ~0ULL;
    }
}

__global__ void k_frontier_deg(const u32 *__restrict__ fr, u64 nf, const u64 *__restrict__ Ap, u64 *__restrict__ deg,
                               u64 *__restrict__ start) {
    u64 t = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    u64 stride = (u64)gridDim.x * blockDim.x;
    for (; t <= nf; t += stride) {
        if (t == nf) { deg[t] = 0; break; }
        u32 u = fr[t];
        u64 s = Ap[u];
        deg[t] = Ap[u + 1] - s;
        start[t] = s;
    }
}

__device__ __forceinline__ u64 bfs_find_le(const u64 *__restrict__ a, u64 lo, u64 hi, u64 target) {
    while (lo < hi) {
        u64 mid = (lo + hi + 1) >> 1;
        if (a[mid] <= target) lo = mid; else hi = mid - 1;
    }
    return lo;
}

// PHASE 0: cand[v] = min(cand[v], u) for unvisited v.  PHASE 1: the winning (u,v) claims v.
template <int PHASE>
__global__ void __launch_bounds__(256)
k_bfs_expand(const u32 *__restrict__ fr, const u64 *__restrict__ cum, const u64 *__restrict__ start, u64 nf, u64 total,
             const u32 *__restrict__ Aj, i64 *__restrict__ level, i64 *__restrict__ parent, u64 *__restrict__ cand,
             i64 next_level, u32 *__restrict__ next, u32 *__restrict__ next_count) {
    __shared__ u64 s_e0, s_e1;
    const u32 tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    u64 lo = (u64)blockIdx.x * BFS_CHUNK, hi = lo + BFS_CHUNK;
    if (hi > total) hi = total;
    if (tid == 0) {
        s_e0 = bfs_find_le(cum, 0, nf - 1, lo);
        s_e1 = bfs_find_le(cum, 0, nf - 1, hi - 1);
    }
    __syncthreads();
    u64 e0 = s_e0, e1 = s_e1;
    for (u64 t0 = lo + (u64)warp * 32; t0 < hi; t0 += 8 * 32) {
        u64 e = bfs_find_le(cum, e0, e1, t0);
        u64 t = t0 + lane;
        if (t < hi) {
            while (e < e1 && cum[e + 1] <= t) e++;
            u32 u = fr[e];
            u32 v = Aj[start[e] + (t - cum[e])];
            if (PHASE == 0) {
                if (level[v] < 0) atomicMin((unsigned long long *)&cand[v], (unsigned long long)u);
            } else {
                // level[v] is only ever written here, by the unique (u,v) with u == cand[v]
                if (cand[v] == (u64)u && level[v] < 0) {
                    level[v] = next_level;
                    if (parent) parent[v] = (i64)u;
                    next[atomicAdd(next_count, 1u)] = v;
                }
            }
        }
    }
}

void bfs_run(const DevCSR &A, u64 src, i64 max_level, i64 *d_level, i64 *d_parent, u64 *edges_traversed) {
    u64 n = A.nrows;
    if (src >= n) throw GrbError(-4, "BFS source out of range");
    DevBuf<u64> cand(n);
    LAUNCH(k_bfs_init, grid_for(n, 256, 148 * 16), 256, 0, d_level, d_parent, cand.ptr, n, src);
    DevBuf<u32> fa(n), fb(n), ncount(1);
    u32 s32 = (u32)src;
    h2d(fa.ptr, &s32, 1);
    u64 nf = 1, edges = 0;
    i64 lvl = 0;
    u32 *cur = fa.ptr, *nxt = fb.ptr;
    DevBuf<u64> cum(n + 1), start(n);
    while (nf > 0 && (max_level < 0 || lvl < max_level)) {
        LAUNCH(k_frontier_deg, grid_for(nf + 1, 256, 148 * 16), 256, 0, cur, nf, A.p.ptr, cum.ptr, start.ptr);
        exclusive_scan_u64(cum.ptr, cum.ptr, nf + 1);
        u64 total = read_scalar(cum.ptr + nf);
        if (total == 0) break;
        edges += total;
        ncount.zero();
        u32 grid = (u32)((total + BFS_CHUNK - 1) / BFS_CHUNK);
        LAUNCH((k_bfs_expand<0>), grid, 256, 0, cur, cum.ptr, start.ptr, nf, total, A.j.ptr, d_level, d_parent, cand.ptr,
               lvl + 1, nxt, ncount.ptr);
        LAUNCH((k_bfs_expand<1>), grid, 256, 0, cur, cum.ptr, start.ptr, nf, total, A.j.ptr, d_level, d_parent, cand.ptr,
               lvl + 1, nxt, ncount.ptr);
        nf = read_scalar(ncount.ptr);
        u32 *t = cur; cur = nxt; nxt = t;
        lvl++;
    }
    if (edges_traversed) *edges_traversed = edges;
}


// ================================================================================================================
// 1-D row-block partitioned BFS (BASELINE config 5).  Rank g owns vertices [row_lo,row_hi) and holds
// their out-edges (a row block of A, rows local-indexed, columns global).  Per level every rank expands the owned
// part of the frontier into an n-bit "discovered" bitmap; the bitmaps are exchanged with ONE all-gather (NCCL, driven
// by the host through torch.distributed) and merged locally; each rank keeps the full visited bitmap (n/8 bytes) and
// the levels of its own vertices.  Parents (deterministic min id) come from a pull over the owned rows of A'.
// ================================================================================================================
__global__ void __launch_bounds__(256)
k_bfs_dist_expand(const u32 *__restrict__ fr, const u64 *__restrict__ cum, const u64 *__restrict__ start, u64 nf, u64 total,
                  const u32 *__restrict__ Aj, const u64 *__restrict__ visited, u32 *__restrict__ disc) {
    __shared__ u64 s_e0, s_e1;
    const u32 tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    u64 lo = (u64)blockIdx.x * BFS_CHUNK, hi = lo + BFS_CHUNK;
    if (hi > total) hi = total;
    if (tid == 0) {
        s_e0 = bfs_find_le(cum, 0, nf - 1, lo);
        s_e1 = bfs_find_le(cum, 0, nf - 1, hi - 1);
    }
    __syncthreads();
    u64 e0 = s_e0, e1 = s_e1;
    for (u64 t0 = lo + (u64)warp * 32; t0 < hi; t0 += 8 * 32) {
        u64 e = bfs_find_le(cum, e0, e1, t0);
        u64 t = t0 + lane;
        if (t < hi) {
            while (e < e1 && cum[e + 1] <= t) e++;
            u32 v = Aj[start[e] + (t - cum[e])];
            if (!((visited[v >> 6] >> (v & 63)) & 1ULL)) atomicOr(&disc[v >> 5], 1u << (v & 31));
        }
    }
}

__global__ void k_frontier_deg_local(const u32 *__restrict__ fr, u64 nf, u64 row_lo, const u64 *__restrict__ Ap,
                                     u64 *__restrict__ deg, u64 *__restrict__ start) {
    u64 t = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    u64 stride = (u64)gridDim.x * blockDim.x;
    for (; t <= nf; t += stride) {
        if (t == nf) { deg[t] = 0; break; }
        u64 u = fr[t] - row_lo;
        u64 s = Ap[u];
        deg[t] = Ap[u + 1] - s;
        start[t] = s;
    }
}

void bfs_dist_expand(const DevCSR &Aloc, u64 row_lo, const u32 *frontier, u64 nf, const u64 *visited, u64 *disc, u64 nwords,
                     u64 *edges_out) {
    CUDA_TRY(cudaMemsetAsync(disc, 0, nwords * sizeof(u64), stream()));
    if (edges_out) *edges_out = 0;
    if (nf == 0) return;
    DevBuf<u64> cum(nf + 1), start(nf);
    LAUNCH(k_frontier_deg_local, grid_for(nf + 1, 256, 148 * 16), 256, 0, frontier, nf, row_lo, Aloc.p.ptr, cum.ptr, start.ptr);
    exclusive_scan_u64(cum.ptr, cum.ptr, nf + 1);
    u64 total = read_scalar(cum.ptr + nf);
    if (edges_out) *edges_out = total;
    if (total == 0) return;
    u32 grid = (u32)((total + BFS_CHUNK - 1) / BFS_CHUNK);
    TimedScope ts(TK_BFS_EXPAND, 4 * total + 20 * nf);
    LAUNCH(k_bfs_dist_expand, grid, 256, 0, frontier, cum.ptr, start.ptr, nf, total, Aloc.j.ptr, visited, (u32 *)disc);
}

// Bottom-up step (direction-optimising BFS): every owned, still-unvisited vertex scans its in-neighbours (a row of the
// owned block of A') for a member of the current frontier bitmap and stops at the first hit.  8 lanes per vertex.
__global__ void __launch_bounds__(256)
k_bfs_dist_pull(const u64 *__restrict__ ATp, const u32 *__restrict__ ATj, u64 nloc, u64 row_lo,
                const u64 *__restrict__ frontier, const u64 *__restrict__ visited, u32 *__restrict__ disc,
                u64 *__restrict__ scanned) {
    const u32 lane8 = threadIdx.x & 7, sub = (threadIdx.x & 31) >> 3;
    const u32 gmask = 0xFFu << (8 * sub);
    u64 group = ((u64)blockIdx.x * blockDim.x + threadIdx.x) >> 3;
    u64 ngroups = ((u64)gridDim.x * blockDim.x) >> 3;
    u64 cnt = 0;
    for (u64 r = group; r < nloc; r += ngroups) {
        u64 v = row_lo + r;
        if ((visited[v >> 6] >> (v & 63)) & 1ULL) continue;       // uniform within the 8-lane group
        u64 s = ATp[r], e = ATp[r + 1];
        bool found = false;
        for (u64 qb = s; qb < e && !found; qb += 8) {
            u64 q = qb + lane8;
            bool hit = false;
            if (q < e) { u32 u = ATj[q]; hit = (frontier[u >> 6] >> (u & 63)) & 1ULL; cnt++; }
            found = __ballot_sync(gmask, hit) & gmask;
        }
        if (found && lane8 == 0) atomicOr(&disc[v >> 5], 1u << (v & 31));
    }
    if (scanned && cnt) atomicAdd((unsigned long long *)scanned, cnt);
}

void bfs_dist_pull(const DevCSR &ATloc, u64 row_lo, const u64 *frontier, const u64 *visited, u64 *disc, u64 nwords,
                   u64 *scanned_out) {
    CUDA_TRY(cudaMemsetAsync(disc, 0, nwords * sizeof(u64), stream()));
    if (scanned_out) *scanned_out = 0;
    if (ATloc.nrows == 0) return;
    DevBuf<u64> sc(1);
    sc.zero();
    {
        TimedScope ts(TK_BFS_EXPAND, 0);
        LAUNCH(k_bfs_dist_pull, grid_for(ATloc.nrows * 8, 256, 148 * 16), 256, 0, ATloc.p.ptr, ATloc.j.ptr, ATloc.nrows, row_lo,
               frontier, visited, (u32 *)disc, sc.ptr);
    }
    if (scanned_out) *scanned_out = read_scalar(sc.ptr);
}

// new = (OR over ranks of the gathered bitmaps) & ~visited ; visited |= new ; owned new vertices get their level and
// join the next local frontier.  counters[0] = next frontier size, counters[1] = global number of new vertices.
__global__ void __launch_bounds__(256)
k_bfs_dist_merge(const u64 *__restrict__ gathered, int P, u64 nwords, u64 *__restrict__ visited, u64 row_lo, u64 row_hi,
                 int *__restrict__ level_local, int lvl, u32 *__restrict__ next, u64 *__restrict__ counters,
                 u64 *__restrict__ frontier_bits) {
    u64 w = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    u64 stride = (u64)gridDim.x * blockDim.x;
    u64 tot = 0;
    for (; w < nwords; w += stride) {
        u64 d = 0;
        for (int g = 0; g < P; g++) d |= gathered[(u64)g * nwords + w];
        u64 old = visited[w];
        u64 nw = d & ~old;
        if (frontier_bits) frontier_bits[w] = nw;   // the next level's frontier as a global bitmap (for the pull step)
        if (!nw) continue;
        visited[w] = old | nw;
        tot += __popcll(nw);
        u64 vb = w << 6;
        if (vb + 64 <= row_lo || vb >= row_hi) continue;
        while (nw) {
            u64 bit = __ffsll((long long)nw) - 1;
            nw &= nw - 1;
            u64 v = vb + bit;
            if (v >= row_lo && v < row_hi) {
                level_local[v - row_lo] = lvl;
                next[atomicAdd((unsigned long long *)&counters[0], 1ULL)] = (u32)v;
            }
        }
    }
    if (tot) atomicAdd((unsigned long long *)&counters[1], tot);
}

void bfs_dist_merge(const u64 *gathered, int P, u64 nwords, u64 *visited, u64 row_lo, u64 row_hi, int *level_local, int lvl,
                    u32 *next, u64 *host_counters, u64 *frontier_bits) {
    DevBuf<u64> cnt(2);
    cnt.zero();
    LAUNCH(k_bfs_dist_merge, grid_for(nwords, 256, 148 * 8), 256, 0, gathered, P, nwords, visited, row_lo, row_hi, level_local, lvl,
           next, cnt.ptr, frontier_bits);
    d2h(host_counters, cnt.ptr, 2);
    sync_stream();
}

// parent(v) = min { u : (u,v) in A, level(u) = level(v) - 1 }: rows of A' are ascending, so the first hit is the minimum
__global__ void k_bfs_dist_parents(const u64 *__restrict__ ATp, const u32 *__restrict__ ATj, u64 nloc, u64 row_lo,
                                   const int *__restrict__ level_full, i64 *__restrict__ parent_local) {
    u64 warp = ((u64)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    u64 nwarps = ((u64)gridDim.x * blockDim.x) >> 5;
    u32 lane = threadIdx.x & 31;
    for (u64 r = warp; r < nloc; r += nwarps) {
        int lv = level_full[row_lo + r];
        i64 par = -1;
        if (lv == 0) par = (i64)(row_lo + r);
        else if (lv > 0) {
            u64 s = ATp[r], e = ATp[r + 1];
            for (u64 q0 = s; q0 < e && par < 0; q0 += 32) {
                u64 q = q0 + lane;
                bool hit = false;
                u32 u = 0;
                if (q < e) { u = ATj[q]; hit = level_full[u] == lv - 1; }
                u32 m = __ballot_sync(0xffffffffu, hit);
                if (m) par = (i64)__shfl_sync(0xffffffffu, u, __ffs(m) - 1);
            }
        }
        if (lane == 0) parent_local[r] = par;
    }
}

void bfs_dist_parents(const DevCSR &ATloc, u64 row_lo, const int *level_full, i64 *parent_local) {
    if (ATloc.nrows == 0) return;
    LAUNCH(k_bfs_dist_parents, grid_for(ATloc.nrows * 32, 256, 148 * 32), 256, 0, ATloc.p.ptr, ATloc.j.ptr, ATloc.nrows, row_lo,
           level_full, parent_local);
}

} // namespace b200
