"""1-D row-block partitioned BFS (BASELINE config 5).
CPU: the partition / direction switch / sparse-vs-dense exchange / termination logic under gloo, world_size 2 and 3, through the
numpy restatement of the library's level loop (falkordb_b200/dist_bfs.py: reference_levels), against the oracle.
GPU: the engine itself (csrc/bfs_do.cu) through the C ABI -- single GPU with every direction policy, the partitioned entry at
world size 1, the dest early exit, and the RMAT block generator -- against the oracle."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _blocks(A, T, lo, hi):
    return (A.p[lo:hi + 1] - A.p[lo], A.j[A.p[lo]:A.p[hi]], T.p[lo:hi + 1] - T.p[lo], T.j[T.p[lo]:T.p[hi]])


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import oracle as orc
    from falkordb_b200.dist_bfs import partition, reference_levels
    A = orc.rmat_csr(10, 8, 3)
    T = orc.transpose(A)
    n = A.nrows
    lo, hi = partition(n, rank, world)
    Ap, Aj, Tp, Tj = _blocks(A, T, lo, hi)
    deg_all = np.diff(A.p)

    def all_gather(a):
        t = torch.from_numpy(a.view(np.int64) if a.dtype == np.uint64 else a.astype(np.int64))
        outs = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(outs, t)
        return [o.numpy().view(np.uint64) if a.dtype == np.uint64 else o.numpy().astype(a.dtype) for o in outs]

    res = []
    srcs = [int(np.nonzero(deg_all)[0][0]), 777, 5]
    for src in srcs:
        for sparse in (True, False):
            lv, par, info = reference_levels(Ap, Aj, Tp, Tj, n, rank, world, deg_all, src, all_gather, sparse_exchange=sparse)
            res.append((src, -1, -1, lo, lv.copy(), par.copy(), info))
    lv, par, info = reference_levels(Ap, Aj, Tp, Tj, n, rank, world, deg_all, srcs[0], all_gather, max_level=2)
    res.append((srcs[0], 2, -1, lo, lv.copy(), par.copy(), info))
    q.put((rank, res))
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_partitioned_bfs_logic_gloo(world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31000 + (os.getpid() * 7 + world) % 2000
    ps = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in ps:
        p.start()
    out = dict(q.get(timeout=240) for _ in ps)
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    import oracle as orc
    A = orc.rmat_csr(10, 8, 3)
    ncase = len(out[0])
    saw = {"td": 0, "bu": 0, "sparse": 0, "dense": 0}
    for k in range(ncase):
        src, max_level = out[0][k][0], out[0][k][1]
        lv = np.concatenate([out[r][k][4] for r in range(world)])[: A.nrows]
        par = np.concatenate([out[r][k][5] for r in range(world)])[: A.nrows]
        wl, wp = orc.bfs(A, src, max_level=max_level)
        assert np.array_equal(lv, wl), f"levels differ for source {src} (case {k})"
        assert np.array_equal(par, wp), f"min-id parents differ for source {src} (case {k})"
        infos = [out[r][k][6] for r in range(world)]
        assert all(i == infos[0] for i in infos), "ranks must take identical control decisions without a collective"
        for key in saw:
            saw[key] += infos[0][key]
    assert saw["td"] and saw["bu"] and saw["sparse"] and saw["dense"], f"every path must have been exercised: {saw}"


# ------------------------------------------------------------------------------------------------ GPU
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
@pytest.mark.parametrize("direction", [0, 1, 2])
def test_direction_optimising_bfs_single_gpu(direction):
    """levels, min-id parents and the Graph500 edge count against the oracle: auto switch, top-down only, bottom-up only"""
    import ctypes as C
    import falkordb_b200 as fb
    import oracle as orc
    from falkordb_b200._lib import lib, check, BfsInfo
    fb.init()
    fb.set_option("bfs_direction", direction)
    try:
        A = orc.rmat_csr(14, 16, 2)
        dA = fb.rmat(14, 16, 2).prepare(True)
        deg = np.diff(A.p)
        n = A.nrows
        for src in (int(np.nonzero(deg)[0][3]), 4097, int(np.argmax(deg))):
            for max_level in (-1, 2):
                lvl, par = np.empty(n, np.int64), np.empty(n, np.int64)
                info = BfsInfo()
                check(lib().B200_bfs_ex(dA.h, src, max_level, -1, lvl.ctypes.data, par.ctypes.data, 0, C.byref(info)))
                wl, wp = orc.bfs(A, src, max_level=max_level)
                assert np.array_equal(lvl, wl) and np.array_equal(par, wp), f"src {src} max_level {max_level}"
                assert info.edges == int(deg[wl >= 0].sum()) and info.depth == int(wl.max())
                if direction == 0 and max_level < 0 and src == int(np.argmax(deg)):
                    assert info.td_levels > 0 and info.bu_levels > 0, "the switch must fire on a scale-14 RMAT sweep"
        # dest early exit (lagraphx_bindings.rs:585-594): levels up to dest's level are final, later ones untouched
        src = int(np.argmax(deg))
        wl, _ = orc.bfs(A, src)
        far = int(np.nonzero(wl == wl.max())[0][0])
        lvl = np.empty(n, np.int64)
        info = BfsInfo()
        check(lib().B200_bfs_ex(dA.h, src, -1, far, lvl.ctypes.data, None, 0, C.byref(info)))
        assert lvl[far] == wl[far] and np.array_equal(lvl[wl <= wl[far]], wl[wl <= wl[far]])
        near = int(np.nonzero(wl == 1)[0][0])
        check(lib().B200_bfs_ex(dA.h, src, -1, near, lvl.ctypes.data, None, 0, C.byref(info)))
        assert info.depth == 1 and lvl[near] == 1 and int((lvl >= 0).sum()) == int(((wl >= 0) & (wl <= 1)).sum())
    finally:
        fb.set_option("bfs_direction", 0)


@pytest.mark.gpu
def test_partitioned_entry_world1_and_lagraph_dest():
    import ctypes as C
    import falkordb_b200 as fb
    import oracle as orc
    from falkordb_b200.dist_bfs import PartitionedBfs
    from falkordb_b200._lib import lib, check, P, obj
    fb.init()
    A = orc.rmat_csr(13, 16, 2)
    deg = np.diff(A.p)
    pb = PartitionedBfs(13, 16, 2, rank=0, world=1)
    try:
        for src in (int(np.nonzero(deg)[0][3]), 4097):
            lv, par, info = pb.run(src)
            wl, wp = orc.bfs(A, src)
            assert np.array_equal(lv, wl) and np.array_equal(par, wp)
            assert info["edges"] == int(deg[wl >= 0].sum()) and info["exchanges"] == 0
    finally:
        pb.close()
    # LAGr_BreadthFirstSearch_Extended with a destination: stops at its level
    L = lib()
    dA = fb.rmat(13, 16, 2)
    G, h = P(), P(dA.h.value)
    dA.h = P()                                                   # LAGraph_New moves the matrix into the graph
    assert L.LAGraph_New(C.byref(G), C.byref(h), 1, None) == 0
    src = int(np.argmax(deg))
    wl, _ = orc.bfs(A, src)
    dest = int(np.nonzero(wl == 2)[0][0])
    lev = P()
    assert L.LAGr_BreadthFirstSearch_Extended(C.byref(lev), None, G, src, -1, dest, False, None) == 0
    nv = C.c_uint64()
    check(L.GrB_Vector_nvals(C.byref(nv), lev))
    I, X = np.empty(nv.value, np.uint64), np.empty(nv.value, np.int64)
    check(L.GrB_Vector_extractTuples_INT64(I.ctypes.data, X.ctypes.data, C.byref(nv), lev))
    got = np.full(A.nrows, -1, np.int64)
    got[I.astype(np.int64)] = X
    assert got[dest] == 2 and np.array_equal(got[wl <= 2], wl[wl <= 2]) and not (got > 2).any()
    L.GrB_Vector_free(C.byref(lev))
    L.LAGraph_Delete(C.byref(G), None)
