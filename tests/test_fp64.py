"""The FP64 corner of the path: GrB_mxv / GrB_vxm over GrB_PLUS_TIMES_SEMIRING_FP64 and GxB_PLUS_SECOND_FP64 (north_star's
"stated fp tolerance for PLUS_TIMES weighted paths"; SURVEY 8d: weights U(0,1) seed 4, rel 1e-12 with a deterministic per-row
summation order) and LAGr_PageRank (algo_procedures.rs:744-752).  Oracle: oracle/grb_oracle.c orc_mxv_fp64 / orc_pagerank,
itself checked against networkx here; the reference's own flow-test expectations (tests/flow/test_pagerank.py) are transcribed."""
import ctypes as C

import numpy as np
import pytest

import falkordb_b200 as fb
import oracle as orc
from falkordb_b200._lib import lib, obj, check, P

REL_TOL = 1e-12          # deterministic summation order


def rel_err(a, b):
    scale = np.maximum(np.abs(b), 1e-300)
    return float(np.max(np.abs(a - b) / scale)) if len(a) else 0.0


def weighted_rmat(scale, seed_w=4):
    A = orc.rmat_csr(scale, 8, 9)
    w = np.random.default_rng(seed_w).random(A.nnz)              # fp64 weights U(0,1), seed 4
    return orc.CSR(A.nrows, A.ncols, A.p, A.j, w.view(np.uint64))


# ------------------------------------------------------------------------------------------ oracle pinned (CPU)
def test_oracle_pagerank_matches_networkx_and_the_reference_flow_expectations():
    nx = pytest.importorskip("networkx")
    names = "ABCDEF"
    idx = {c: i for i, c in enumerate(names)}
    edges = [("A", "B"), ("B", "C"), ("C", "F"), ("F", "E"), ("E", "D"), ("D", "A"), ("E", "B")]   # tests/flow/test_pagerank.py:56-71
    A = orc.build_matrix(6, 6, [idx[a] for a, _ in edges], [idx[b] for _, b in edges])
    r, it = orc.pagerank(A)
    assert abs(r.sum() - 1.0) < 1e-4 and (r > 0).all() and r[idx["B"]] == r.max()                  # test_pagerank.py:91-104
    S = orc.build_matrix(2, 2, [0], [1])                                                          # S1 -> S2 (test_pagerank.py:116-150)
    rs, _ = orc.pagerank(S)
    assert rs[1] > rs[0] and abs(rs.sum() - 1.0) < 1e-4
    B = orc.rmat_csr(10, 8, 3)                                    # sinks and sources galore
    rb, itb = orc.pagerank(B, tol=1e-13, itermax=2000)
    G = nx.DiGraph()
    G.add_nodes_from(range(B.nrows))
    rr, cc, _ = B.tuples()
    G.add_edges_from(zip(rr.tolist(), cc.tolist()))
    want = nx.pagerank(G, alpha=0.85, tol=1e-15, max_iter=5000)
    assert max(abs(rb[k] - v) for k, v in want.items()) < 1e-12


def test_oracle_mxv_fp64_matches_scipy():
    A = weighted_rmat(9)
    x = np.random.default_rng(1).random(A.ncols)
    y, yp = orc.mxv_fp64(A, x, use_values=True)
    import scipy.sparse as sp
    M = sp.csr_matrix((A.x.view(np.float64), A.j.astype(np.int64), A.p), shape=(A.nrows, A.ncols))
    assert rel_err(y[yp], (M @ x)[yp]) < 1e-13
    assert np.array_equal(yp, np.diff(A.p) > 0)


# ------------------------------------------------------------------------------------------ CUDA path
def _vec_from(L, arr, present=None):
    v = P()
    check(L.GrB_Vector_new(C.byref(v), obj("GrB_FP64"), len(arr)))
    for i in (range(len(arr)) if present is None else np.nonzero(present)[0]):
        check(L.GrB_Vector_setElement_FP64(v, float(arr[i]), int(i)))
    return v


def _vec_to(L, v, n):
    nv = C.c_uint64(n)
    I, X = np.empty(n, np.uint64), np.empty(n, np.float64)
    check(L.GrB_Vector_extractTuples_FP64(I.ctypes.data, X.ctypes.data, C.byref(nv), v))
    return I[: nv.value].astype(np.int64), X[: nv.value]


def _fp64_matrix(L, A):
    m = P()
    check(L.GrB_Matrix_new(C.byref(m), obj("GrB_FP64"), A.nrows, A.ncols))
    rows = np.repeat(np.arange(A.nrows, dtype=np.uint64), np.diff(A.p))
    cols = A.j.astype(np.uint64)
    vals = A.x.view(np.float64).copy()
    check(L.GrB_Matrix_build_FP64(m, rows.ctypes.data, cols.ctypes.data, vals.ctypes.data, A.nnz, None))
    return m


@pytest.mark.gpu
def test_plus_times_and_plus_second_mxv_vxm_against_the_oracle():
    fb.init()
    L = lib()
    A = weighted_rmat(12)
    AT = orc.transpose(A)
    n = A.nrows
    rng = np.random.default_rng(5)
    x = rng.random(n)
    present = rng.random(n) < 0.3
    dA = _fp64_matrix(L, A)
    nv = C.c_uint64(A.nnz)
    Xb = np.empty(A.nnz, np.float64)
    check(L.GrB_Matrix_extractTuples_FP64(None, None, Xb.ctypes.data, C.byref(nv), dA))
    assert np.array_equal(Xb, A.x.view(np.float64)), "FP64 values must survive build -> extract bit for bit"
    cases = [("GrB_PLUS_TIMES_SEMIRING_FP64", True), ("GxB_PLUS_SECOND_FP64", False)]
    for sr, use_values in cases:
        for pres in (None, present):
            u = _vec_from(L, x, pres)
            w = P()
            check(L.GrB_Vector_new(C.byref(w), obj("GrB_FP64"), n))
            # mxv: w = A*u
            check(L.GrB_mxv(w, None, None, obj(sr), dA, u, None))
            I, X = _vec_to(L, w, n)
            wy, wp = orc.mxv_fp64(A, x, use_values, pres)
            assert np.array_equal(I, np.nonzero(wp)[0]), f"{sr}: pattern of A*u"
            assert rel_err(X, wy[wp]) < REL_TOL, f"{sr}: A*u rel err {rel_err(X, wy[wp])}"
            # vxm: w' = u'*A  ==  A'*u
            check(L.GrB_vxm(w, None, None, obj(sr), u, dA, None))
            I, X = _vec_to(L, w, n)
            wy, wp = orc.mxv_fp64(AT, x, use_values, pres)
            assert np.array_equal(I, np.nonzero(wp)[0]) and rel_err(X, wy[wp]) < REL_TOL, f"{sr}: u*A"
            # mxv with T0 == vxm
            check(L.GrB_mxv(w, None, None, obj(sr), dA, u, obj("GrB_DESC_T0")))
            I2, X2 = _vec_to(L, w, n)
            assert np.array_equal(I2, I) and np.array_equal(X2, X), "A'*u through the descriptor must equal u*A bit for bit"
            # accum PLUS: w += A*u, run twice from the result
            check(L.GrB_mxv(w, None, obj("GrB_PLUS_FP64"), obj(sr), dA, u, None))
            I3, X3 = _vec_to(L, w, n)
            ay, ap = orc.mxv_fp64(A, x, use_values, pres)
            full = np.zeros(n)
            full[I] = X
            keep = np.zeros(n, bool)
            keep[I] = True
            keep |= ap
            full = full + np.where(ap, ay, 0.0)
            assert np.array_equal(I3, np.nonzero(keep)[0]) and rel_err(X3, full[keep]) < REL_TOL
            L.GrB_Vector_free(C.byref(u)); L.GrB_Vector_free(C.byref(w))
    # run-to-run determinism: identical bits
    u = _vec_from(L, x)
    w = P()
    check(L.GrB_Vector_new(C.byref(w), obj("GrB_FP64"), n))
    outs = []
    for _ in range(3):
        check(L.GrB_mxv(w, None, None, obj("GrB_PLUS_TIMES_SEMIRING_FP64"), dA, u, None))
        outs.append(_vec_to(L, w, n)[1].copy())
    assert np.array_equal(outs[0], outs[1]) and np.array_equal(outs[1], outs[2])
    L.GrB_Matrix_free(C.byref(dA))


def _pagerank_dev(L, A_dev_handle, n, damping=0.85, tol=1e-4, itermax=100):
    G, h = P(), P(A_dev_handle.value)
    assert L.LAGraph_New(C.byref(G), C.byref(h), 1, None) == 0
    assert L.LAGraph_Cached_AT(G, None) == 0 and L.LAGraph_Cached_OutDegree(G, None) == 0     # algo_procedures.rs:748-749
    cen, it = P(), C.c_int(0)
    assert L.LAGr_PageRank(C.byref(cen), C.byref(it), G, damping, tol, itermax, None) == 0
    I, X = _vec_to(L, cen, n)
    L.GrB_Vector_free(C.byref(cen))
    L.LAGraph_Delete(C.byref(G), None)
    assert np.array_equal(I, np.arange(n))
    return X, it.value


@pytest.mark.gpu
def test_lagr_pagerank_against_the_oracle_and_the_reference_flow_tests():
    fb.init()
    L = lib()
    from falkordb_b200.grb import Matrix

    def to_dev(c):
        return Matrix.import_csr(c.nrows, c.ncols, c.p.astype(np.uint64), c.j, None, bool)
    names = "ABCDEF"
    idx = {c: i for i, c in enumerate(names)}
    edges = [("A", "B"), ("B", "C"), ("C", "F"), ("F", "E"), ("E", "D"), ("D", "A"), ("E", "B")]
    A6 = orc.build_matrix(6, 6, [idx[a] for a, _ in edges], [idx[b] for _, b in edges])
    for A, tol, itermax in ((A6, 1e-4, 100), (orc.rmat_csr(12, 8, 3), 1e-4, 100), (orc.rmat_csr(13, 16, 7), 1e-9, 200)):
        m = to_dev(A)
        h = P(m.h.value)
        m.h = P()                                    # the matrix moves into the LAGraph graph (lagraph_bindings.rs:175)
        got, it = _pagerank_dev(L, h, A.nrows, 0.85, tol, itermax)
        # the C entry takes float damping / tol (lagraph_bindings.rs:554-555): the oracle gets the same rounded values
        want, wit = orc.pagerank(A, float(np.float32(0.85)), float(np.float32(tol)), itermax)
        assert it == wit, f"iteration count {it} vs oracle {wit}"
        assert rel_err(got, want) < 1e-9, f"scores rel err {rel_err(got, want)}"
        assert abs(got.sum() - 1.0) < 1e-4 and (got > 0).all()                      # tests/flow/test_pagerank.py:94-104
    g6, _ = _pagerank_dev(L, _detach(to_dev(A6)), 6)
    assert g6[idx["B"]] == g6.max()                  # B has two in-edges: highest score (test_pagerank.py:91-96)


def _detach(m):
    h = P(m.h.value)
    m.h = P()
    return h
