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
