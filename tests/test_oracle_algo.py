"""CPU checks of the oracle's restatements of the LAGraph kernels behind algo.labelPropagation and algo.WCC, against brute force and
against what the reference's own flow tests assert (tests/flow/test_cdlp.py:83-178, tests/flow/test_wcc.py): the oracle is what
the GPU tests compare LAGraph_cdlp / LAGr_ConnectedComponents with."""
from collections import Counter

import numpy as np

import oracle as orc


def sym(n, pairs):
    s = np.array([a for a, b in pairs] + [b for a, b in pairs], np.int64)
    d = np.array([b for a, b in pairs] + [a for a, b in pairs], np.int64)
    return orc.build_matrix(n, n, s, d)


def cdlp_brute(A, itermax):
    n = A.nrows
    L = list(range(n))
    rounds = 0
    while rounds < itermax and A.nnz:
        new = list(L)
        for v in range(n):
            nb = A.j[A.p[v]:A.p[v + 1]]
            if len(nb):
                c = Counter(L[int(u)] for u in nb)
                top = max(c.values())
                new[v] = min(l for l, k in c.items() if k == top)
        rounds += 1
        same = new == L
        L = new
        if same:
            break
    return np.array(L, np.int64), rounds


def test_cdlp_matches_brute_force_on_random_symmetric_graphs():
    rng = np.random.default_rng(5)
    for n, m, itermax in ((1, 0, 10), (7, 0, 3), (12, 20, 10), (60, 150, 10), (60, 150, 1), (200, 260, 4), (300, 3000, 10)):
        pairs = [(int(a), int(b)) for a, b in rng.integers(0, n, (m, 2))]      # self-edges and repeats included
        A = sym(n, pairs) if m else orc.build_matrix(n, n, np.zeros(0, np.int64), np.zeros(0, np.int64))
        got, r = orc.cdlp(A, itermax)
        want, rw = cdlp_brute(A, itermax)
        assert np.array_equal(got, want) and r == rw, (n, m, itermax)


def test_cdlp_reference_flow_test_shapes():
    """tests/flow/test_cdlp.py:83-178: three fully connected triples come out as exactly three communities, one per triple;
    the same with the four triples of the label / relationship filter cases (:180-330)."""
    tri = lambda a: [(a, a + 1), (a, a + 2), (a + 1, a + 2)]
    for k in (3, 4):
        A = sym(3 * k, sum((tri(3 * c) for c in range(k)), []))
        L, rounds = orc.cdlp(A, 10)
        assert len(set(L.tolist())) == k
        for c in range(k):
            assert L[3 * c] == L[3 * c + 1] == L[3 * c + 2] == 3 * c        # min-label ties: the triple's smallest id
        assert rounds <= 10


def test_wcc_matches_union_find():
    rng = np.random.default_rng(9)
    for n, m in ((1, 0), (50, 30), (400, 350), (400, 2000)):
        pairs = [(int(a), int(b)) for a, b in rng.integers(0, n, (m, 2))]
        A = sym(n, pairs) if m else orc.build_matrix(n, n, np.zeros(0, np.int64), np.zeros(0, np.int64))
        parent = list(range(n))

        def find(x):
            while parent[x] != x:
                parent[x] = parent[parent[x]]
                x = parent[x]
            return x
        for a, b in pairs:
            ra, rb = find(a), find(b)
            if ra != rb:
                parent[max(ra, rb)] = min(ra, rb)
        want = np.array([find(v) for v in range(n)], np.int64)
        assert np.array_equal(orc.wcc(A), want)


def test_masked_product_dot_form_equals_saxpy_form_then_mask():
    """The oracle evaluates C<M> = A*B in dot form (oracle/grb_oracle.c: mxm_masked_dot; nothing outside the mask is ever formed --
    the saxpy form materialised ~1e11 unmasked entries for config 4 at full size).  It must equal the unmasked saxpy product with
    the mask applied afterwards, entry for entry, count the same flops (sum of deg_B(k) over A's entries: SURVEY 8d), and agree
    with scipy -- on the triangle pattern L*L<L>, on a rectangular product with an unrelated mask, and with empty rows."""
    import scipy.sparse as sp
    rng = np.random.default_rng(4)
    A0 = orc.rmat_csr(12, 8, 3)
    U = orc.ewise_add(A0, orc.transpose(A0))
    n = U.nrows
    rows = np.repeat(np.arange(n, dtype=np.int64), np.diff(U.p))
    keep = U.j.astype(np.int64) < rows
    L = orc.build_matrix(n, n, rows[keep], U.j[keep])
    cases = [(L, L, L)]
    A = orc.build_matrix(300, 500, rng.integers(0, 300, 4000), rng.integers(0, 500, 4000))
    B = orc.build_matrix(500, 400, rng.integers(0, 500, 6000), rng.integers(0, 400, 6000))
    M = orc.build_matrix(300, 400, rng.integers(0, 150, 9000), rng.integers(0, 400, 9000))      # rows 150.. of the mask are empty
    cases.append((A, B, M))
    for A_, B_, M_ in cases:
        got, fl = orc.mxm(A_, B_, M_, 1, return_flops=True)
        full = orc.mxm(A_, B_)
        want = orc.mask_assign(None, full, M_, comp=False, structural=True, replace=True)
        assert np.array_equal(got.p, want.p) and np.array_equal(got.j, want.j)
        assert fl == int(np.diff(B_.p)[A_.j].sum())
        S = (A_.to_scipy().astype(np.int64) @ B_.to_scipy().astype(np.int64)).multiply(M_.to_scipy().astype(np.int64)).tocsr()
        S.eliminate_zeros()
        S.sort_indices()
        assert np.array_equal(got.p, S.indptr) and np.array_equal(got.j, S.indices)
