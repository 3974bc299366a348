"""Oracle (oracle/grb_oracle.c) vs scipy.sparse -- an independent implementation -- and vs
algebraic identities (the reference holds no direct known answer for raw mxm)."""
import numpy as np
import pytest
import scipy.sparse as sp

import oracle as orc


def rand_csr(rng, nrows, ncols, density, values=False):
    m = sp.random(nrows, ncols, density=density, format="csr", random_state=rng,
                  data_rvs=lambda k: rng.integers(0, 5, k))
    m.data = m.data.astype(np.int64)
    return orc.CSR.from_scipy(m, values=values)


def pat(m):
    m = m.tocsr().copy()
    m.data[:] = 1
    m.eliminate_zeros()
    m.sort_indices()
    return m


def same_pattern(c, s):
    s = pat(s)
    return np.array_equal(c.p, s.indptr) and np.array_equal(c.j, s.indices.astype(np.uint32))


@pytest.mark.parametrize("seed", range(6))
def test_mxm_matches_scipy(seed):
    rng = np.random.default_rng(seed)
    n, k, m = rng.integers(1, 300, 3)
    A = rand_csr(rng, n, k, 0.05)
    B = rand_csr(rng, k, m, 0.05)
    Cc, flops = orc.mxm(A, B, return_flops=True)
    assert same_pattern(Cc, pat(A.to_scipy()) @ pat(B.to_scipy()))
    degB = np.diff(B.p)
    assert flops == int(degB[A.j].sum())


def test_mxm_masked_matches_set_algebra():
    rng = np.random.default_rng(7)
    A = rand_csr(rng, 120, 90, 0.08)
    B = rand_csr(rng, 90, 150, 0.08)
    M = rand_csr(rng, 120, 150, 0.2)
    full = orc.mxm(A, B).tuple_set()
    ms = M.tuple_set()
    assert orc.mxm(A, B, M, 1).tuple_set() == full & ms
    assert orc.mxm(A, B, M, 2).tuple_set() == full - ms
    # C<M> u C<!M> == C (SURVEY 8c identity ii)
    assert orc.mxm(A, B, M, 1).tuple_set() | orc.mxm(A, B, M, 2).tuple_set() == full


def test_mxm_associative_and_values_ignored():
    rng = np.random.default_rng(3)
    F = rand_csr(rng, 40, 200, 0.02)
    A = rand_csr(rng, 200, 200, 0.03, values=True)   # u64 operand: values never read
    B = rand_csr(rng, 200, 200, 0.03)
    left = orc.mxm(orc.mxm(F, A), B)
    right = orc.mxm(F, orc.mxm(A, B))
    assert left == right
    assert orc.mxm(F, A) == orc.mxm(F, orc.pattern(A))


def test_mxm_dimension_mismatch():
    rng = np.random.default_rng(0)
    with pytest.raises(ValueError):
        orc.mxm(rand_csr(rng, 5, 6, 0.5), rand_csr(rng, 7, 5, 0.5))


def test_empty_and_ragged():
    E = orc.CSR.empty(10, 10)
    rng = np.random.default_rng(1)
    A = rand_csr(rng, 10, 10, 0.3)
    assert orc.mxm(E, A).nnz == 0 and orc.mxm(A, E).nnz == 0
    assert orc.ewise_add(E, A) == A
    assert orc.ewise_mult(E, A).nnz == 0
    assert orc.transpose(E).nnz == 0


@pytest.mark.parametrize("seed", range(4))
def test_ewise_and_transpose_match_scipy(seed):
    rng = np.random.default_rng(100 + seed)
    n, m = rng.integers(1, 200, 2)
    A = rand_csr(rng, n, m, 0.1, values=True)
    B = rand_csr(rng, n, m, 0.1, values=True)
    assert same_pattern(orc.ewise_add(A, B), pat(A.to_scipy()) + pat(B.to_scipy()))
    assert same_pattern(orc.ewise_mult(A, B), pat(A.to_scipy()).multiply(pat(B.to_scipy())))
    # SECOND: B's value wins on overlap, single-side copied (matrix.rs:277-281)
    U = orc.ewise_add(A, B, keep_values=True)
    da = {(r, c): v for r, c, v in A.tuple_set()}
    db = {(r, c): v for r, c, v in B.tuple_set()}
    da.update(db)
    assert U.tuple_set() == {(r, c, v) for (r, c), v in da.items()}
    T = orc.transpose(A)
    assert T.tuple_set() == {(c, r, v) for r, c, v in A.tuple_set()}
    assert orc.transpose(T) == A   # involution


def test_mask_assign_semantics():
    rng = np.random.default_rng(11)
    Cold = rand_csr(rng, 50, 50, 0.1)
    T = rand_csr(rng, 50, 50, 0.1)
    M = rand_csr(rng, 50, 50, 0.3, values=True)  # values in 0..4: valued mask differs from structural
    c, t = Cold.tuple_set(), T.tuple_set()
    ms = {(r, cc) for r, cc, v in M.tuple_set()}
    mv = {(r, cc) for r, cc, v in M.tuple_set() if v != 0}
    assert ms != mv
    for comp in (False, True):
        for structural in (False, True):
            for replace in (False, True):
                mk = ms if structural else mv
                allpos = c | t
                inmask = {e for e in allpos if ((e in mk) != comp)}
                want = (t & inmask) | (set() if replace else (c - inmask))
                got = orc.mask_assign(Cold, T, M, comp, structural, replace).tuple_set()
                assert got == want, (comp, structural, replace)
                # accum = ANY: Z = Cold u T inside the mask
                wanta = ((c | t) & inmask) | (set() if replace else (c - inmask))
                gota = orc.mask_assign(Cold, T, M, comp, structural, replace, accum=True).tuple_set()
                assert gota == wanta
    # no mask, no accum: C = T
    assert orc.mask_assign(Cold, T) == T


def test_build_collapses_duplicates_first_wins():
    m = orc.build_matrix(8, 8, [1, 3, 1, 3, 1], [2, 4, 2, 4, 2])
    assert m.nnz == 2 and m.tuple_set() == {(1, 2), (3, 4)}
    v = orc.build_matrix(8, 8, [1, 3, 1], [2, 4, 2], [10, 20, 30])
    assert v.tuple_set() == {(1, 2, 10), (3, 4, 20)}
    with pytest.raises(IndexError):
        orc.build_matrix(4, 4, [4], [0])


def test_bfs_levels_match_scipy_and_parents_are_min():
    from scipy.sparse.csgraph import breadth_first_order, shortest_path
    rng = np.random.default_rng(5)
    A = rand_csr(rng, 300, 300, 0.01)
    lvl, par = orc.bfs(A, 0)
    d = shortest_path(pat(A.to_scipy()), method="D", unweighted=True, indices=0)
    want = np.where(np.isinf(d), -1, d).astype(np.int64)
    assert np.array_equal(lvl, want)
    assert par[0] == 0
    S = pat(A.to_scipy()).tocsc()
    for v in np.nonzero(lvl > 0)[0]:
        preds = S.indices[S.indptr[v]:S.indptr[v + 1]]
        ok = preds[lvl[preds] == lvl[v] - 1]
        assert par[v] == ok.min()
    l2, _ = orc.bfs(A, 0, max_level=2)
    assert np.array_equal(l2, np.where(want <= 2, want, -1))


def test_rmat_deterministic_and_clean():
    a = orc.rmat_csr(10, 16, seed=1)
    b = orc.rmat_csr(10, 16, seed=1)
    c = orc.rmat_csr(10, 16, seed=2)
    assert a == b and not (a == c)
    r, cc, _ = a.tuples()
    assert not np.any(r == cc)                       # no self loops
    assert len(a.tuple_set()) == a.nnz               # no duplicates
    assert np.all(np.diff(a.p) >= 0) and a.nnz > 8 * 1024
    # skew: RMAT max degree far above mean
    assert np.diff(a.p).max() > 20 * a.nnz / a.nrows


def test_delta_lmxm_equals_materialised_merge():
    """matrix.rs:1305-1402: delta_lmxm == F * ((m u dp) \\ dm) when dm subset of m, dp disjoint m."""
    rng = np.random.default_rng(21)
    n = 150
    m = rand_csr(rng, n, n, 0.05)
    dp_raw = rand_csr(rng, n, n, 0.01)
    dp = orc.mask_assign(None, dp_raw, m, comp=True, structural=True, replace=True)  # dp \ m
    sel = rand_csr(rng, n, n, 0.3)
    dm = orc.ewise_mult(m, sel)                                                      # dm subset m
    F = rand_csr(rng, 30, n, 0.03)
    eff = orc.mask_assign(None, orc.ewise_add(m, dp), dm, comp=True, structural=True, replace=True)
    got = orc.delta_lmxm(F, m, dp, dm)
    # NOTE the reference formula masks out every (i,j) reachable through a deleted edge even if
    # another live edge reaches j; it equals F*eff only when no j has both.  Check the formula
    # itself, and the equality on the clean snapshot.
    mk = orc.mxm(F, dm)
    want = (orc.mxm(F, m).tuple_set() - mk.tuple_set()) | orc.mxm(F, dp).tuple_set()
    assert got.tuple_set() == want
    assert orc.delta_lmxm(F, eff, orc.CSR.empty(n, n), orc.CSR.empty(n, n)) == orc.mxm(F, eff)


def test_bfs_reference_flow_goldens():
    """known answers of the reference's own query-level BFS tests (tests/flow/test_bfs.py:9-25, 63-175): the graph
    (a)-[:E1]->(b)-[:E1]->(c), (b)-[:E2]->(d)-[:E1]->(e); `nodes` = every vertex with level >= 1 (0-hop excluded)"""
    a, b, c, d, e = range(5)
    E1 = [(a, b), (b, c), (d, e)]
    E2 = [(b, d)]

    def adj(edges):
        return orc.build_matrix(5, 5, [s for s, _ in edges], [t for _, t in edges])

    def reached(A, src, max_level=-1):
        lvl, par = orc.bfs(A, src, max_level)
        assert lvl[src] == 0 and par[src] == src                      # algo_procedures.rs:1098-1148 convention
        return sorted(int(v) for v in np.nonzero(lvl >= 1)[0]), lvl, par

    ALL = adj(E1 + E2)
    nodes, lvl, par = reached(ALL, a)                                  # test01: algo.BFS(a, -1, NULL)
    assert nodes == [b, c, d, e] and list(lvl) == [0, 1, 2, 2, 3]
    assert [int(par[v]) for v in (b, c, d, e)] == [a, b, b, d]          # `edges` = the tree edge into each node
    A1 = adj(E1)
    assert reached(A1, a)[0] == [b, c]                                 # test02: restricted to E1
    want3 = {a: [b, c], b: [c], d: [e]}                                # test03: all sources, E1; others yield nothing
    for src in range(5):
        assert reached(A1, src)[0] == want3.get(src, [])
    assert reached(ALL, a, 1)[0] == [b]                                # test04: max depth 1
    want5 = {a: [b], b: [c, d], d: [e]}                                # test05: all sources, depth 1
    for src in range(5):
        assert reached(ALL, src, 1)[0] == want5.get(src, [])


def test_variable_length_reference_flow_goldens():
    """tests/flow/test_variable_length_traversals.py:3-6, 51-64: on the chain A->B->C->D, MATCH (a)-[*]->(b) returns 6 rows
    (A reaches 3, B 2, C 1) in either direction.  Paths are unique there, so the pair set equals the level-wise reachability
    loop falkordb_b200.multi_source_reach runs (C<!R,replace> = F*A ; R = R u F), restated here on the oracle."""
    A = orc.build_matrix(4, 4, [0, 1, 2], [1, 2, 3])
    for M_, want in ((A, {(0, 1), (0, 2), (0, 3), (1, 2), (1, 3), (2, 3)}), (orc.transpose(A), {(1, 0), (2, 0), (3, 0), (2, 1), (3, 1), (3, 2)})):
        F = orc.build_matrix(4, 4, range(4), range(4))
        R = orc.build_matrix(4, 4, [], [])
        while True:
            F = orc.mxm(F, M_, R, mask_mode=2) if R.nnz else orc.mxm(F, M_)
            if F.nnz == 0:
                break
            R = orc.ewise_add(R, F)
        assert R.tuple_set() == want and R.nnz == 6
    one_hop = orc.mxm(orc.build_matrix(4, 4, range(4), range(4)), A)      # test01: (a)-[e]->(b): AB, BC, CD
    assert one_hop.tuple_set() == {(0, 1), (1, 2), (2, 3)}
