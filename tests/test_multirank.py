"""N>1 host logic on CPU (gloo, world_size 2): the path shards the frontier rows across ranks with the adjacency
replicated and NO data-path collective; the only communication is the whole-job reduction of the
timing (max over ranks) and work counters (sum)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import bench
    import oracle as orc
    A = orc.rmat_csr(10, 8, 1)                      # replicated on every rank
    deg = np.diff(A.p)
    batches = bench.pick_sources(deg, 3, 16, seed=1, rank=rank)
    flops = 0
    rows = []
    for b in batches:                               # each rank expands only its own sources
        F, fl = bench.cpu_chain(orc, A, b, 2)
        flops += fl
        rows.append((b.copy(), F.p.copy(), F.j.copy()))
    (tmax,), (fsum, nsum) = bench.reduce_over_ranks([10.0 + rank], [flops, sum(int(r[1][-1]) for r in rows)], "cpu")
    out.put((rank, tmax, fsum, nsum, flops, [b.tolist() for b in batches]))
    dist.destroy_process_group()


def test_sharded_sources_and_whole_job_reduction():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    ps = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    res = sorted(q.get(timeout=120) for _ in ps)
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, t0, f0, n0, own0, b0), (r1, t1, f1, n1, own1, b1) = res
    assert t0 == t1 == 11.0                         # max over ranks
    assert f0 == f1 == own0 + own1                  # work is summed
    assert n0 == n1
    assert b0 != b1                                 # ranks draw different (seeded) source batches
    # row independence: expanding the union of both ranks' sources equals the concatenation of the per-rank results
    import bench
    import oracle as orc
    A = orc.rmat_csr(10, 8, 1)
    both = np.array(b0[0] + b1[0], dtype=np.uint64)
    F, _ = bench.cpu_chain(orc, A, both, 2)
    Fa, _ = bench.cpu_chain(orc, A, np.array(b0[0], dtype=np.uint64), 2)
    Fb, _ = bench.cpu_chain(orc, A, np.array(b1[0], dtype=np.uint64), 2)
    assert np.array_equal(F.j, np.concatenate([Fa.j, Fb.j]))
    assert np.array_equal(F.p, np.concatenate([Fa.p, Fb.p[1:] + Fa.p[-1]]))


def test_triangle_row_blocks_balance_the_intersection_work():
    """BASELINE config 4 at N > 1 (bench.py: run_triangles): row blocks of L = tril(A u A') with equal intersection work, every rank
    computing the same cuts from the replicated L.  The cuts must tile the rows, balance the kernel's work model to a few percent
    where equal-row blocks are off by integer factors, and -- rows of the masked product being independent -- the per-block results
    must concatenate to the whole product."""
    import bench
    import oracle as orc
    A = orc.rmat_csr(14, 16, 1)
    U = orc.ewise_add(A, orc.transpose(A))
    n = U.nrows
    rows = np.repeat(np.arange(n, dtype=np.int64), np.diff(U.p))
    keep = U.j.astype(np.int64) < rows
    L = orc.build_matrix(n, n, rows[keep], U.j[keep])
    full = orc.mxm(L, L, L, 1)
    for world in (2, 4, 8):
        cuts, work = bench.triangle_row_cuts(L.p, L.j, world)
        assert cuts[0] == 0 and cuts[-1] == n and all(a <= b for a, b in zip(cuts, cuts[1:])) and len(cuts) == world + 1
        mean = sum(work) / world
        assert max(work) <= 1.05 * mean, (world, work)
        # the same work model over equal-row blocks: what round 1 did
        eq = [n * g // world for g in range(world + 1)]
        deg = np.diff(L.p)
        w_entry = np.minimum(deg[L.j], deg[np.repeat(np.arange(n), deg)]) + 8
        csum = np.concatenate([[0], np.cumsum(w_entry)])
        eq_work = [int(csum[L.p[eq[g + 1]]] - csum[L.p[eq[g]]]) for g in range(world)]
        assert max(eq_work) > 1.5 * (sum(eq_work) / world) > 0, "equal-row blocks of tril(L) are badly skewed (the reason for the cuts)"
        parts_p, parts_j = [np.zeros(1, np.int64)], []
        for g in range(world):
            lo, hi = cuts[g], cuts[g + 1]
            Lb = orc.CSR(hi - lo, n, L.p[lo:hi + 1] - L.p[lo], L.j[L.p[lo]:L.p[hi]])
            Cb = orc.mxm(Lb, L, Lb, 1)
            parts_p.append(Cb.p[1:] + parts_p[-1][-1])
            parts_j.append(Cb.j)
        assert np.array_equal(np.concatenate(parts_p), full.p) and np.array_equal(np.concatenate(parts_j), full.j)
