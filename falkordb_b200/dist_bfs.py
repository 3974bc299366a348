"""1-D row-block partitioned BFS (BASELINE config 5): host-side glue around B200_bfs_partitioned.

Rank g owns vertices [lo, hi): the out-edges (row block of A) and the in-edges (row block of A') of those vertices.  The level
loop, the direction switch and the exchange (NCCL all-gather on the library's stream: n-bit bitmaps, or sentinel-padded vertex
lists while the frontier is sparse) all live in the library (csrc/bfs_do.cu); this module only
  * builds the row blocks of the synthetic RMAT graph (B200_Matrix_rmat_block regenerates the counter-based edge stream, so the
    union of the blocks is exactly the single-GPU matrix),
  * creates the NCCL communicator (rank 0 makes the 128-byte id, torch.distributed ships it),
  * and offers `reference_levels`, a numpy restatement of the same level loop over any all_gather callable: the gloo /
    world_size-2 CPU tests run the partition, switch and exchange logic through it.
"""
import ctypes as C

import numpy as np

ALPHA, BETA = 14, 24          # Beamer's switch constants, as in csrc/bfs_do.cu


def partition(n, rank, world):
    """Contiguous blocks of ceil(n / world) rounded up to 64 vertices (bitmap words never straddle two owners)."""
    block = -(-n // world)
    block = (block + 63) // 64 * 64
    lo = min(n, rank * block)
    hi = min(n, lo + block)
    return lo, hi


# ------------------------------------------------------------------------------------------------ numpy restatement
def _bits_of(idx, nwords):
    b = np.zeros(nwords, np.uint64)
    idx = np.asarray(idx, dtype=np.int64)
    np.bitwise_or.at(b, idx >> 6, np.uint64(1) << (idx & 63).astype(np.uint64))
    return b


def _members(bits, n):
    return np.nonzero(np.unpackbits(bits.view(np.uint8), bitorder="little")[:n])[0]


def reference_levels(Ap, Aj, ATp, ATj, n, rank, world, deg_all, src, all_gather, max_level=-1, dest=-1, sparse_exchange=True):
    """bfs_do's level loop on host arrays.  Ap/Aj: CSR of the owned row block of A (local rows, global columns); ATp/ATj: the
    same rows of A'.  all_gather(np.ndarray) -> list of every rank's array (equal shapes).  Returns (level, parent) of the
    owned vertices (int64, -1 = unreached) and a dict of what happened."""
    lo, hi = partition(n, rank, world)
    block = partition(n, 0, world)[1] if world > 1 else n
    nwl = block // 64 if world > 1 else (n + 63) // 64
    nw = nwl * world
    visited = _bits_of([src], nw)
    frontier = visited.copy()
    level = np.full(hi - lo, -1, np.int64)
    parent = np.full(hi - lo, -1, np.int64)
    if lo <= src < hi:
        level[src - lo], parent[src - lo] = 0, src
    total_edges = int(np.asarray(deg_all, dtype=np.int64).sum())
    nf, mf = 1, int(deg_all[src])
    explored = mf
    info = {"td": 0, "bu": 0, "sparse": 0, "dense": 0}
    bottom_up, lvl = False, 0

    def first_frontier_member(r, fr):
        nb = ATj[ATp[r]:ATp[r + 1]].astype(np.int64)
        hit = ((fr[nb >> 6] >> (nb & 63).astype(np.uint64)) & np.uint64(1)).astype(bool)
        return int(nb[np.argmax(hit)]) if hit.any() else -1

    while nf > 0 and (max_level < 0 or lvl < max_level) and dest != src:
        unexplored = max(0, total_edges - explored)
        if not bottom_up:
            if mf > unexplored // ALPHA:
                bottom_up = True
        elif nf < n // BETA:
            bottom_up = False
        lvl += 1
        if not bottom_up:
            mine = [v for v in _members(frontier, n) if lo <= v < hi]
            tgt = np.concatenate([Aj[Ap[v - lo]:Ap[v - lo + 1]] for v in mine]).astype(np.int64) if mine else np.zeros(0, np.int64)
            tgt = tgt[((visited[tgt >> 6] >> (tgt & 63).astype(np.uint64)) & np.uint64(1)) == 0]
            disc = _bits_of(tgt, nw)
            if world > 1 and sparse_exchange and mf < n // 32:
                bound = max(1, mf)
                lst = np.full(bound, 0xFFFFFFFF, np.uint32)
                mem = _members(disc, n)
                lst[:len(mem)] = mem
                allv = np.concatenate(all_gather(lst))
                allv = allv[allv != 0xFFFFFFFF].astype(np.int64)
                new = _bits_of(allv, nw) & ~visited
                info["sparse"] += 1
            else:
                parts = all_gather(disc) if world > 1 else [disc]
                new = np.bitwise_or.reduce(np.stack(parts), axis=0) & ~visited
                info["dense"] += world > 1
            for v in _members(new, n):
                if lo <= v < hi:
                    level[v - lo] = lvl
                    parent[v - lo] = first_frontier_member(v - lo, frontier)
            info["td"] += 1
        else:
            new = np.zeros(nw, np.uint64)
            found = []
            for r in range(hi - lo):
                v = lo + r
                if (visited[v >> 6] >> np.uint64(v & 63)) & np.uint64(1):
                    continue
                p = first_frontier_member(r, frontier)
                if p >= 0:
                    level[r], parent[r] = lvl, p
                    found.append(v)
            new |= _bits_of(found, nw)
            if world > 1:
                w0 = rank * nwl
                parts = all_gather(new[w0:w0 + nwl].copy())
                new = np.concatenate(parts)
            info["bu"] += 1
        mem = _members(new, n)
        nf, mf = len(mem), int(np.asarray(deg_all, dtype=np.int64)[mem].sum())
        visited |= new
        explored += mf
        if nf == 0:
            lvl -= 1
            break
        frontier = new
        if dest >= 0 and dest in set(mem.tolist()):
            break
    info["depth"] = lvl
    return level, parent, info


# ------------------------------------------------------------------------------------------------ GPU
class PartitionedBfs:
    """Row blocks of the RMAT graph on this rank's GPU + the NCCL communicator; `run` is one collective BFS."""

    def __init__(self, scale, edge_factor, seed, rank, world, broadcast_bytes=None):
        """broadcast_bytes(bytes or None) -> bytes: ships rank 0's 128-byte NCCL id to every rank (torch.distributed in
        bench.py and the tests); not needed for world == 1."""
        from ._lib import lib, check, P
        self.L, self.check = lib(), check
        self.n = 1 << scale
        self.rank, self.world = rank, world
        self.lo, self.hi = partition(self.n, rank, world)
        self.A, self.AT = P(), P()
        check(self.L.B200_Matrix_rmat_block(C.byref(self.A), scale, edge_factor, seed, self.lo, self.hi, 0))
        check(self.L.B200_Matrix_rmat_block(C.byref(self.AT), scale, edge_factor, seed, self.lo, self.hi, 1))
        self.comm = P()
        ident = None
        if world > 1:
            buf = (C.c_uint8 * 128)()
            if rank == 0:
                check(self.L.B200_comm_unique_id(buf))
            ident = broadcast_bytes(bytes(buf) if rank == 0 else None)
            buf = (C.c_uint8 * 128).from_buffer_copy(ident)
            check(self.L.B200_comm_init(C.byref(self.comm), rank, world, buf))
        else:
            check(self.L.B200_comm_init(C.byref(self.comm), 0, 1, None))

    def local_degrees(self):
        p = np.empty(self.hi - self.lo + 1, np.uint64)
        self.check(self.L.B200_Matrix_export_CSR(self.A, p.ctypes.data, None, None, 0))
        return np.diff(p.astype(np.int64))

    def run(self, src, max_level=-1, dest=-1, want_parents=True, on_device=False):
        """Returns (level_local int64[hi-lo], parent_local or None, info dict).  Collective: every rank calls it.
        on_device: the results stay in HBM (torch tensors) instead of being copied to host arrays."""
        from ._lib import BfsInfo
        nloc = self.hi - self.lo
        info = BfsInfo()
        if on_device:
            import torch
            level = torch.empty(max(1, nloc), dtype=torch.int64, device="cuda")
            parent = torch.empty(max(1, nloc), dtype=torch.int64, device="cuda") if want_parents else None
            lp, pp, loc = level.data_ptr(), (None if parent is None else parent.data_ptr()), 1
        else:
            level = np.empty(max(1, nloc), np.int64)
            parent = np.empty(max(1, nloc), np.int64) if want_parents else None
            lp, pp, loc = level.ctypes.data, (None if parent is None else parent.ctypes.data), 0
        self.check(self.L.B200_bfs_partitioned(self.A, self.AT, self.n, self.lo, self.comm, int(src), int(max_level), int(dest),
                                               lp, pp, loc, C.byref(info)))
        return level[:nloc], (None if parent is None else parent[:nloc]), info.as_dict()

    def close(self):
        if self.comm.value:
            self.L.B200_comm_free(C.byref(self.comm))
        for h in (self.A, self.AT):
            if h.value:
                self.L.GrB_Matrix_free(C.byref(h))
