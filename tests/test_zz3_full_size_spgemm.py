"""Bit-exact parity at the sizes BASELINE.json's configs 2, 3 and 4 name, against the CPU oracle on the same CSR inputs (digest
comparison: see tests/test_zz1_full_size_chain_bfs.py and tests/fullsize_util.py).  Config 4 = masked triangles on RMAT-24
(fused masked SpGEMM), config 3 = LDBC SF10-shaped 3-relationship chain, config 2 = one GrB_mxm on RMAT-22 through the row-wise
SpGEMM with its multi-wave heavy-row path.  Never run on hardware in the form committed here."""
import os
import time

import numpy as np
import pytest

import falkordb_b200 as fb
import oracle as orc
from falkordb_b200.grb import Matrix, Descriptor
from fullsize_util import need, rmat_both, same_digest, rows_of, dev_of, defaults, _GRAPHS

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _defaults():
    yield from defaults()


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
