"""Bit-exact parity at the sizes BASELINE.json's configs name, against the CPU oracle on the same CSR inputs.
Multi-GB results are compared through an order-sensitive digest computed on both sides (B200_Matrix_digest on the device,
orc_digest on the host: nvals, sum mix(row << 32 | col), sum mix(key + GOLD * CSR position)) plus exact flops; BFS levels and
parents are compared element for element.  The graph is generated on the device and exported once, so both sides read the
same CSR.  Each test skips only when the box lacks the memory it needs (free HBM / host RAM probe), never by default.

The module sorts last on purpose (pytest -x): these are the longest tests, and the only ones not yet run on hardware in the form
committed here -- the session that wrote them lost its GPU access when the first version of the config-4 oracle (saxpy form: every
unmasked row materialised, ~1e11 entries on RMAT-24) exhausted a test box's RAM.  The oracle now evaluates masked products in dot
form (oracle/grb_oracle.c: mxm_masked_dot, a few GB), every test states its host-memory need up front (`need`), and
tests/conftest.py runs a resident-set watchdog that kills the test process long before the host is in danger."""
import os
import time

import numpy as np
import pytest

import falkordb_b200 as fb
import oracle as orc
from falkordb_b200.grb import Matrix, Descriptor

pytestmark = pytest.mark.gpu


# The whole `pytest -m gpu` run has a wall-clock limit where it is driven from (20 minutes in round 1's record).  A full-size test
# that would START later than this many seconds into the session skips, loudly, rather than take the run over the limit and
# turn every earlier pass into a timeout; on the 128-thread box the module needs about five minutes, so this is a backstop.
START_BY_S = float(os.environ.get("B200_FULLSIZE_START_BY_S", "840"))


def need(hbm_gb, host_gb):
    import torch
    from conftest import SESSION_T0
    late = time.time() - SESSION_T0
    if late > START_BY_S:
        pytest.skip(f"{late:.0f} s into the session (> {START_BY_S:.0f}): not starting a multi-minute test; raise B200_FULLSIZE_START_BY_S")
    free, _ = torch.cuda.mem_get_info()
    if free < hbm_gb * 2 ** 30:
        pytest.skip(f"needs {hbm_gb} GiB of free HBM, {free / 2 ** 30:.0f} available")
    try:
        import psutil
        avail = psutil.virtual_memory().available
    except Exception:
        avail = os.sysconf("SC_AVPHYS_PAGES") * os.sysconf("SC_PAGE_SIZE")
    if avail < host_gb * 2 ** 30:
        pytest.skip(f"needs {host_gb} GiB of host memory, {avail / 2 ** 30:.0f} available")


@pytest.fixture(autouse=True)
def _defaults():
    fb.init()
    for k, v in (("bits_mode", -1), ("pull_mode", -1), ("small_cap", 4096), ("bitmap_budget", 2 << 30), ("bits_min_flops", 1 << 22)):
        fb.set_option(k, v)
    orc.lib().orc_set_num_threads(len(os.sched_getaffinity(0)))
    yield
    fb.lib().B200_pool_trim()


_GRAPHS = {}


def rmat_both(scale, seed=1):
    """(device matrix, oracle CSR of the same graph): generated on the device, exported once."""
    key = (scale, seed)
    if key not in _GRAPHS:
        _GRAPHS.clear()                      # one resident graph at a time
        A = fb.rmat(scale, 16, seed)
        p, j, _ = A.export_csr()
        _GRAPHS[key] = (A, orc.CSR(1 << scale, 1 << scale, p.astype(np.int64), j))
    return _GRAPHS[key]


def same_digest(dev, want_digest, what):
    got = dev.digest()
    assert int(got[0]) == int(want_digest[0]), f"{what}: nvals {int(got[0])} != {int(want_digest[0])}"
    assert np.array_equal(got, want_digest), f"{what}: digests differ (same nvals): {got} vs {want_digest}"


def rows_of(Ao, rows):
    """F = A(rows, :) as an oracle CSR (len(rows) x n)"""
    deg = np.diff(Ao.p)[rows]
    p = np.zeros(len(rows) + 1, np.int64)
    np.cumsum(deg, out=p[1:])
    j = np.empty(int(p[-1]), np.uint32)
    for k, r in enumerate(rows):             # rows is sorted; slices are contiguous copies
        j[p[k]:p[k + 1]] = Ao.j[Ao.p[r]:Ao.p[r + 1]]
    return orc.CSR(len(rows), Ao.ncols, p, j)


def dev_of(c):
    return Matrix.import_csr(c.nrows, c.ncols, c.p.astype(np.uint64), c.j, None, bool)


# ------------------------------------------------------------------------------------------ headline: 3-hop chain, RMAT-24
@pytest.mark.parametrize("nsrc", [512, 1000])
def test_chain_rmat24_all_rows(nsrc):
    """the bench workload, every row: 3-hop F*A*A*A from `nsrc` random non-isolated sources on RMAT-24 (W = 8 / 16 word columns:
    the lane-split pull, long-row segments, the multi-word materialise)"""
    need(40, 60)
    A, Ao = rmat_both(24)
    A.prepare(True)
    deg = np.diff(Ao.p)
    rng = np.random.default_rng(1000003)
    src = rng.choice(np.nonzero(deg > 0)[0], size=nsrc, replace=False).astype(np.uint64)
    _, flops, dg, busy = orc.chain(Ao, src, 3, keep=False)
    F = Matrix(nsrc, Ao.nrows, bool)
    F.build(np.arange(nsrc, dtype=np.uint64), src)
    fl = 0
    for _ in range(3):
        F.lmxm(A)
        fl += fb.get_stat("last_flops")
    assert fb.get_stat("last_path") == 3, "the last hop must take the pull direction at this size"
    F.wait()
    assert fl == flops, "flops (edges traversed) differ from the oracle's"
    same_digest(F, dg, f"3-hop chain, {nsrc} sources, RMAT-24")
    # the bitmap hand-off of the same result: row populations must match the CSR row lengths
    p = np.empty(nsrc + 1, np.uint64)
    fb.check(fb.lib().B200_Matrix_export_CSR(F.h, p.ctypes.data, None, None, 0))
    G = Matrix(nsrc, Ao.nrows, bool)
    G.build(np.arange(nsrc, dtype=np.uint64), src)
    for _ in range(3):
        G.lmxm(A)
    wpr = (Ao.nrows + 63) // 64
    bm = np.empty((nsrc, wpr), np.uint64)
    G.export_bitmap(bm)
    assert np.array_equal(np.bitwise_count(bm).sum(axis=1).astype(np.int64), np.diff(p.astype(np.int64)))


# ------------------------------------------------------------------------------------------ config 5: BFS, RMAT-26
def test_config5_bfs_rmat26_levels_and_parents():
    """BASELINE config 5's graph on one GPU: BFS level (bit-exact) and min-id parent vectors from 3 random sources with out-edges
    on RMAT scale-26 (n = 67.1 M, ~1.05e9 edges) against the oracle; algo.BFS's conventions (algo_procedures.rs:1098-1148)."""
    need(90, 40)
    A, Ao = rmat_both(26)
    A.prepare(True)               # the transpose mirror: B200_bfs then runs the direction-optimising engine (bfs_do.cu)
    deg = np.diff(Ao.p)
    rng = np.random.default_rng(3)
    srcs = rng.choice(np.nonzero(deg > 0)[0], size=3, replace=False)
    for s in srcs:
        lvl, par, edges = fb.bfs(A, int(s))
        wl, wp = orc.bfs(Ao, int(s))
        assert np.array_equal(lvl, wl), f"levels differ from the oracle (source {s})"
        assert np.array_equal(par, wp), f"min-id parents differ from the oracle (source {s})"
        assert edges == int(deg[wl >= 0].sum()), "edges traversed (Graph500 convention) differ"
    _GRAPHS.clear()


# ------------------------------------------------------------------------------------------ config 4: masked triangles, RMAT-24
def test_config4_masked_triangles_rmat24():
    """BASELINE config 4 at its stated size: C<L, struct, replace> = L*L over ANY_PAIR with L = tril(A u A') of RMAT-24: which
    edges close a wedge.  The oracle evaluates the product row by row and drops what the mask excludes."""
    need(40, 24)          # host: L (1.3 GB) + L' + one byte per mask entry + the result; the dot-form oracle forms nothing unmasked
    import ctypes as C
    from falkordb_b200._lib import lib, check, P as VP
    _GRAPHS.clear()
    scale = 24
    n = 1 << scale
    h = VP()
    check(lib().B200_Matrix_rmat_block(C.byref(h), scale, 16, 1, 0, n, 2))
    L = Matrix(0, 0, bool, _handle=h)
    p, j, _ = L.export_csr()
    Lo = orc.CSR(n, n, p.astype(np.int64), j)
    rows = np.repeat(np.arange(n, dtype=np.int64), np.diff(Lo.p))
    assert np.all(Lo.j.astype(np.int64) < rows), "L must be strictly lower triangular"
    del rows
    want, flops = orc.mxm(Lo, Lo, Lo, 1, return_flops=True)
    dg = orc.digest(want)
    nnz_want = want.nnz
    del want
    fb.set_option("bits_mode", 0)
    Cm = Matrix(n, n, bool)
    Cm.mxm(L, L, L, Descriptor.RS)
    assert fb.get_stat("last_flops") == flops
    assert Cm.nvals() == nnz_want
    same_digest(Cm, dg, f"config 4: nnz(L) {Lo.nnz}, flops {flops}, closed wedges {nnz_want}")


# ------------------------------------------------------------------------------------------ config 3: LDBC SF10-shaped chain
def ldbc_sf10(seed=10):
    """SF10-shaped synthetic social graph with the cardinalities SURVEY 8(d) lists (external LDBC-spec figures, NOT from the
    reference): 66,000 Person, 7,500,000 Post, 16,080 Tag; KNOWS ~1.9 M directed (power-law, symmetric), hasCreator^T one
    Person -> Post edge per post (7.5 M), hasTag ~10 M Post -> Tag.  All matrices n x n with label ranges (graph.rs:1191)."""
    P, Q, T = 66_000, 7_500_000, 16_080
    rng = np.random.default_rng(seed)
    n = P + Q + T
    deg = np.minimum(P - 1, (rng.pareto(1.5, P) * 5 + 1).astype(np.int64))
    deg = (deg * (950_000 / deg.sum())).astype(np.int64) + 1            # ~0.95 M undirected -> ~1.9 M directed
    ks = np.repeat(np.arange(P), deg)
    kd = rng.integers(0, P, len(ks))
    keep = ks != kd
    knows = orc.build_matrix(n, n, np.concatenate([ks[keep], kd[keep]]), np.concatenate([kd[keep], ks[keep]]))
    creator = np.minimum(P - 1, (rng.pareto(1.2, Q) * P / 20).astype(np.int64))
    created = orc.build_matrix(n, n, creator, P + np.arange(Q))
    ntag = rng.choice([1, 2], size=Q, p=[2 / 3, 1 / 3])                  # mean 4/3 -> ~10 M hasTag edges
    ps = np.repeat(P + np.arange(Q), ntag)
    tg = P + Q + np.minimum(T - 1, (rng.pareto(1.1, len(ps)) * T / 50).astype(np.int64))
    hastag = orc.build_matrix(n, n, ps, tg)
    return P, Q, T, n, knows, created, hastag


def test_config3_ldbc_sf10_shaped_chain():
    """BASELINE config 3 at SF10 cardinalities: F = all Persons (66,000 rows); F <- F*KNOWS*CREATED*HASTAG.  The operator would
    feed <= 1024-row batches (batch.rs:81); here the whole frontier goes through GrB_mxm at once (row-wise SpGEMM), and the
    first 1024 Persons also ride the frontier bit-matrix path; both against the oracle."""
    need(30, 40)
    P, Q, T, n, knows, created, hastag = ldbc_sf10()
    counts = f"Person {P}, Post {Q}, Tag {T}, KNOWS {knows.nnz}, hasCreator^T {created.nnz}, hasTag {hastag.nnz}"
    assert 1_700_000 < knows.nnz < 2_100_000 and created.nnz == Q and 9_000_000 < hastag.nnz < 11_000_000, counts
    dK, dC, dH = dev_of(knows), dev_of(created), dev_of(hastag)
    for rows in (P, 1024):
        want = orc.build_matrix(rows, n, np.arange(rows), np.arange(rows))
        F = Matrix(rows, n, bool)
        F.build(np.arange(rows, dtype=np.uint64), np.arange(rows, dtype=np.uint64))
        fl_want = 0
        for M_, dM in ((knows, dK), (created, dC), (hastag, dH)):
            want, f1 = orc.mxm(want, M_, return_flops=True)
            F.lmxm(dM)
            assert fb.get_stat("last_flops") == f1, counts
            fl_want += f1
        F.wait()
        assert want.nnz > 0 and int(want.j.min()) >= P + Q, "the chain must land in the tag range"
        same_digest(F, orc.digest(want), f"config 3 ({rows} rows): {counts}")


# ------------------------------------------------------------------------------------------ config 2: single mxm, RMAT-22
@pytest.mark.parametrize("variant", ["frontier_2e18", "F_eq_A_row_block"])
def test_config2_single_mxm_rmat22(variant):
    """BASELINE config 2 at its stated size: one GrB_mxm over ANY_PAIR on RMAT scale-22 through the row-wise SpGEMM.
    frontier_2e18: F = the rows of A for a random 2^18-vertex frontier (seed 2).  F_eq_A_row_block: F = A, evaluated for one
    2^17-row block -- the whole product A*A has ~7.8e10 entries (312 GB of column indices; measured growth x7.8 per two
    scales: 1.28e9 at scale 18), more than one GPU's HBM, so the F = A variant is checked block-wise."""
    need(60, 120)         # host: the oracle's rows (~5e9 entries at 2^18 frontier rows: 20 GB) three times over (arena, result, numpy copy)
    A, Ao = rmat_both(22)
    n = Ao.nrows
    rng = np.random.default_rng(2)
    if variant == "frontier_2e18":
        rows = np.sort(rng.choice(n, size=1 << 18, replace=False))
    else:
        lo = int(rng.integers(0, n - (1 << 17)))
        rows = np.arange(lo, lo + (1 << 17))
    F = rows_of(Ao, rows)
    want, flops = orc.mxm(F, Ao, return_flops=True)
    dg = orc.digest(want)
    nnz_want = want.nnz
    del want
    C_ = Matrix(F.nrows, n, bool)
    C_.mxm(dev_of(F), A)
    C_.wait()
    assert fb.get_stat("last_flops") == flops == int(np.diff(Ao.p)[F.j].sum())
    assert C_.nvals() == nnz_want
    same_digest(C_, dg, f"config 2 {variant}: {F.nrows} x {n}, flops {flops}, nnz(C) {nnz_want}")
