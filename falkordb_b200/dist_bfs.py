"""1-D row-block partitioned BFS over `torch.distributed` (BASELINE config 5).

Rank g owns vertices [lo, hi) and the out-edges of those vertices (a row block of A).  One level =
  expand : owned part of the frontier  ->  n-bit "discovered" bitmap           (CUDA kernel, bfs.cu)
  gather : ONE all-gather of the bitmaps over NCCL / NVLink                     (torch.distributed)
  merge  : OR the P bitmaps, drop visited, assign levels to owned vertices,
           emit the next owned frontier                                         (CUDA kernel, bfs.cu)
Every rank keeps the full visited bitmap (n/8 bytes); the loop ends when the merged bitmap is empty, which every rank
sees identically, so no extra termination collective is needed.  Parents (deterministic minimum id, the rule of the
single-GPU kernel and of the oracle) come from one all-gather of the int32 levels and a pull over the owned rows of A'.

The level loop is backend-agnostic: `GpuBackend` drives libb200grb.so on device tensors, tests substitute a numpy
backend under gloo to check the partition / exchange logic on CPU.
"""
import ctypes as C

import numpy as np


def partition(n, rank, world):
    """Contiguous blocks, multiples of 64 vertices so bitmap words never straddle two owners."""
    block = -(-n // world)
    block = (block + 63) // 64 * 64
    lo = min(n, rank * block)
    hi = min(n, lo + block)
    return lo, hi


def run_levels(backend, n, rank, world, src, all_gather, max_level=-1, pull_threshold=1 / 64):
    """Level loop.  Returns (levels of owned vertices as int32, out-edges of the owned reached vertices, depth).
    Direction-optimising: when the previous level discovered more than `pull_threshold * n` vertices the next level runs
    bottom-up (owned unvisited vertices look for a frontier in-neighbour) if the backend offers it."""
    lo, hi = partition(n, rank, world)
    backend.reset(src)
    nf = 1 if lo <= src < hi else 0
    lvl, total_new = 0, 1
    can_pull = hasattr(backend, "expand_pull") and backend.can_pull()
    while max_level < 0 or lvl < max_level:
        if can_pull and total_new > pull_threshold * n:
            backend.expand_pull()
        else:
            backend.expand(nf)
        gathered = all_gather(backend.disc())
        nf, total_new = backend.merge(gathered, lvl + 1)
        if total_new == 0:
            break
        lvl += 1
    return backend.levels(), backend.reached_edges(), lvl


class GpuBackend:
    """Device-resident state in torch tensors; kernels through the C ABI on raw device pointers."""

    def __init__(self, scale, edge_factor, seed, rank, world, need_parents=True):
        import torch
        from ._lib import lib, check, P
        self.torch, self.L, self.check = torch, lib(), check
        self.n = 1 << scale
        self.rank, self.world = rank, world
        self.lo, self.hi = partition(self.n, rank, world)
        self.nwords = (self.n + 63) // 64
        h = P()
        check(self.L.B200_Matrix_rmat_block(C.byref(h), scale, edge_factor, seed, self.lo, self.hi, 0))
        self.A = h
        self.AT = None
        if need_parents:
            ht = P()
            check(self.L.B200_Matrix_rmat_block(C.byref(ht), scale, edge_factor, seed, self.lo, self.hi, 1))
            self.AT = ht
        dev = torch.device("cuda", torch.cuda.current_device())
        nloc = max(1, self.hi - self.lo)
        self.visited = torch.zeros(self.nwords, dtype=torch.int64, device=dev)
        self._disc = torch.zeros(self.nwords, dtype=torch.int64, device=dev)
        self.frontier_bits = torch.zeros(self.nwords, dtype=torch.int64, device=dev)
        self.level = torch.full((nloc,), -1, dtype=torch.int32, device=dev)
        p = np.empty(self.hi - self.lo + 1, np.uint64)
        check(self.L.B200_Matrix_export_CSR(self.A, p.ctypes.data, None, None, 0))
        self.deg = torch.from_numpy(np.diff(p.astype(np.int64))).to(dev)
        self.fa = torch.zeros(nloc, dtype=torch.int32, device=dev)
        self.fb = torch.zeros(nloc, dtype=torch.int32, device=dev)
        self.cnt = np.zeros(2, np.uint64)

    def reset(self, src):
        t = self.torch
        self.visited.zero_()
        self.level.fill_(-1)
        w, b = src >> 6, src & 63
        self.visited[w] = (1 << b) if b < 63 else -(1 << 63)
        if self.lo <= src < self.hi:
            self.level[src - self.lo] = 0
            self.fa[0] = src if src < (1 << 31) else src - (1 << 32)
        t.cuda.synchronize()

    def expand(self, nf):
        e = C.c_uint64(0)
        self.check(self.L.B200_bfs_dist_expand(self.A, self.lo, self.fa.data_ptr(), nf, self.visited.data_ptr(),
                                               self._disc.data_ptr(), self.nwords, C.byref(e)))
        return e.value

    def can_pull(self):
        return self.AT is not None

    def expand_pull(self):
        sc = C.c_uint64(0)
        self.check(self.L.B200_bfs_dist_pull(self.AT, self.lo, self.frontier_bits.data_ptr(), self.visited.data_ptr(),
                                             self._disc.data_ptr(), self.nwords, C.byref(sc)))
        return sc.value

    def reached_edges(self):
        """Graph500 edge count: out-edges of the owned vertices that were reached."""
        lv = self.level[: self.hi - self.lo]
        return int(self.deg[lv >= 0].sum().item()) if self.hi > self.lo else 0

    def disc(self):
        return self._disc

    def merge(self, gathered, lvl):
        self.torch.cuda.synchronize()                     # NCCL ran on torch's stream; the library has its own
        self.check(self.L.B200_bfs_dist_merge(gathered.data_ptr(), self.world, self.nwords, self.visited.data_ptr(), self.lo,
                                              self.hi, self.level.data_ptr(), lvl, self.fb.data_ptr(), self.cnt.ctypes.data,
                                              self.frontier_bits.data_ptr()))
        self.fa, self.fb = self.fb, self.fa
        return int(self.cnt[0]), int(self.cnt[1])

    def levels(self):
        return self.level[: self.hi - self.lo]

    def parents(self, level_full):
        """level_full: int32[n] on the device (all-gathered).  Returns int64 parents of the owned vertices."""
        par = self.torch.full((max(1, self.hi - self.lo),), -1, dtype=self.torch.int64, device=self.level.device)
        self.torch.cuda.synchronize()
        self.check(self.L.B200_bfs_dist_parents(self.AT, self.lo, level_full.data_ptr(), par.data_ptr()))
        return par[: self.hi - self.lo]

    def close(self):
        for h in (self.A, self.AT):
            if h is not None and h.value:
                self.L.GrB_Matrix_free(C.byref(h))


def bfs_gpu(backend, src, max_level=-1, want_parents=True):
    """Whole distributed BFS on the GPU backend.  Returns (level_local, parent_local or None, edges_local, depth)."""
    import torch
    import torch.distributed as dist
    world = backend.world

    def all_gather(disc):
        if world == 1:
            return disc
        out = torch.empty(world * disc.numel(), dtype=disc.dtype, device=disc.device)
        dist.all_gather_into_tensor(out, disc)
        return out

    lv, edges, depth = run_levels(backend, backend.n, backend.rank, world, src, all_gather, max_level)
    par = None
    if want_parents:
        block = partition(backend.n, 0, world)[1]
        if world == 1:
            full = lv
        else:
            pad = torch.full((block,), -1, dtype=torch.int32, device=lv.device)
            pad[: lv.numel()] = lv
            full = torch.empty(world * block, dtype=torch.int32, device=lv.device)
            dist.all_gather_into_tensor(full, pad)
        par = backend.parents(full.contiguous())
    return lv, par, edges, depth
