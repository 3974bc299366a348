"""1-D row-partitioned BFS (falkordb_b200/dist_bfs.py).
CPU: the partition / exchange / termination logic under gloo, world_size 2, with a numpy stand-in for the three device
kernels.  GPU: the real kernels through the C ABI at world size 1, and the RMAT block generator, against the oracle."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


class NumpyBackend:
    """Same contract as dist_bfs.GpuBackend, host arrays (test double for the CUDA kernels)."""

    def __init__(self, A, rank, world):
        from falkordb_b200.dist_bfs import partition
        self.n, self.rank, self.world = A.nrows, rank, world
        self.lo, self.hi = partition(self.n, rank, world)
        self.p = A.p[self.lo:self.hi + 1] - A.p[self.lo]
        self.j = A.j[A.p[self.lo]:A.p[self.hi]]
        self.nwords = (self.n + 63) // 64

    def reset(self, src):
        self.visited = np.zeros(self.nwords, np.uint64)
        self.visited[src >> 6] |= np.uint64(1) << np.uint64(src & 63)
        self.level = np.full(self.hi - self.lo, -1, np.int32)
        self.front = np.array([src], np.int64) if self.lo <= src < self.hi else np.zeros(0, np.int64)
        if self.lo <= src < self.hi:
            self.level[src - self.lo] = 0

    def expand(self, nf):
        assert nf == len(self.front)
        self._disc = np.zeros(self.nwords, np.uint64)
        edges = 0
        for u in self.front:
            nb = self.j[self.p[u - self.lo]:self.p[u - self.lo + 1]].astype(np.int64)
            edges += len(nb)
            seen = (self.visited[nb >> 6] >> (nb & 63).astype(np.uint64)) & np.uint64(1)
            nb = nb[seen == 0]
            np.bitwise_or.at(self._disc, nb >> 6, np.uint64(1) << (nb & 63).astype(np.uint64))
        return edges

    def disc(self):
        return torch.from_numpy(self._disc.view(np.int64))

    def merge(self, gathered, lvl):
        g = gathered.numpy().view(np.uint64).reshape(-1, self.nwords)
        new = np.bitwise_or.reduce(g, axis=0) & ~self.visited
        self.visited |= new
        bits = np.unpackbits(new.view(np.uint8), bitorder="little")[: self.n]
        v = np.nonzero(bits)[0]
        own = v[(v >= self.lo) & (v < self.hi)]
        self.level[own - self.lo] = lvl
        self.front = own
        return len(own), len(v)

    def levels(self):
        return self.level

    def reached_edges(self):
        return int(np.diff(self.p)[self.level >= 0].sum())


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import oracle as orc
    from falkordb_b200.dist_bfs import run_levels
    A = orc.rmat_csr(10, 8, 3)
    be = NumpyBackend(A, rank, world)

    def all_gather(t):
        outs = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(outs, t)
        return torch.cat(outs)

    res = []
    for src in (int(np.nonzero(np.diff(A.p))[0][0]), 777, 5):
        lv, edges, depth = run_levels(be, A.nrows, rank, world, src, all_gather)
        res.append((src, be.lo, lv.copy(), edges, depth))
    l2, _, d2 = run_levels(be, A.nrows, rank, world, res[0][0], all_gather, max_level=2)
    res.append((res[0][0], be.lo, l2.copy(), -1, d2))
    q.put((rank, res))
    dist.destroy_process_group()


def test_partitioned_bfs_logic_world2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31000 + os.getpid() % 2000
    ps = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    out = dict(q.get(timeout=180) for _ in ps)
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    import oracle as orc
    A = orc.rmat_csr(10, 8, 3)
    deg = np.diff(A.p)
    for k in range(4):
        src = out[0][k][0]
        lv = np.concatenate([out[0][k][2], out[1][k][2]])
        want, _ = orc.bfs(A, src, max_level=2 if k == 3 else -1, want_parent=False)
        assert np.array_equal(lv, want.astype(np.int32)), f"levels differ for source {src}"
        if k < 3:   # every reached vertex is expanded exactly once, by its owner: Graph500's edge count
            assert out[0][k][3] + out[1][k][3] == int(deg[want >= 0].sum())
            assert out[0][k][4] == out[1][k][4]       # both ranks stop at the same level without a collective


@pytest.mark.gpu
def test_rmat_blocks_tile_the_full_matrix():
    import ctypes as C
    import falkordb_b200 as fb
    import oracle as orc
    from falkordb_b200._lib import lib, check, P
    from falkordb_b200.grb import Matrix
    fb.init()
    A = orc.rmat_csr(11, 16, 5)
    T = orc.transpose(A)
    S = orc.ewise_add(A, T)
    r, c, _ = S.tuples()
    Lo = orc.build_matrix(S.nrows, S.ncols, r[r > c], c[r > c])     # tril(A u A'), the masked-SpGEMM operand (config 4)
    for by_col, ref in ((0, A), (1, T), (2, Lo)):
        for lo, hi in ((0, 704), (704, 2048), (0, 2048)):
            h = P()
            check(lib().B200_Matrix_rmat_block(C.byref(h), 11, 16, 5, lo, hi, by_col))
            m = Matrix(0, 0, bool, _handle=h)
            p, j, _ = m.export_csr()
            assert np.array_equal(p, ref.p[lo:hi + 1] - ref.p[lo]) and np.array_equal(j, ref.j[ref.p[lo]:ref.p[hi]])


@pytest.mark.gpu
def test_partitioned_bfs_kernels_world1_match_oracle():
    import falkordb_b200 as fb
    import oracle as orc
    from falkordb_b200.dist_bfs import GpuBackend, bfs_gpu
    fb.init()
    A = orc.rmat_csr(13, 16, 2)
    be = GpuBackend(13, 16, 2, rank=0, world=1)
    try:
        for src in (int(np.nonzero(np.diff(A.p))[0][3]), 4097):
            lv, par, edges, depth = bfs_gpu(be, src)
            wl, wp = orc.bfs(A, src)
            assert np.array_equal(lv.cpu().numpy(), wl.astype(np.int32))
            assert np.array_equal(par.cpu().numpy(), wp)
            assert edges == int(np.diff(A.p)[wl >= 0].sum())
    finally:
        be.close()
