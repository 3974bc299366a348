"""GPU tests written after round 2's last hardware session (gpurun closed before they could run on a B200).  The module sorts after
every hardware-verified one and after the most certain full-size tests (tests/test_zz1_*), before the heavy SpGEMM configs
(tests/test_zz3_*) and the ordered-frontier observers (tests/test_zz4_*): `pytest -x` stops at the first failure, so the order is
most-certain-first.  What is here:
  * LAGr_ConnectedComponents and LAGraph_cdlp (algo.cu) against the oracle's restatements (scipy / numpy; tests/test_oracle_algo.py
    checks those against brute force on the CPU)
  * the output batch re-pack of the host mirror (cond_traverse.hpp: repack)
  * the trail enumerator over a Tensor (cond_var_len_traverse.hpp; its DFS logic alone is CPU-tested in test_host_varlen.py)
  * B200_Tensor_bulk_build (GRAPH.BULK into an empty tensor) against a numpy model and against the host mirror's insert loop
  * the Tensor's C-compatible RDB form after device-side mutations (its host-only half is tests/test_host_serial.py)"""
import ctypes as C

import numpy as np
import pytest

import falkordb_b200 as fb
import oracle as orc
from falkordb_b200._lib import lib, obj, check, P, U64
from falkordb_b200.grb import Matrix, Descriptor
from test_gpu_parity import to_dev, assert_same, bitmap_of, diag_csr
from test_host_tensor import run as run_host_test

pytestmark = pytest.mark.gpu


def test_lagr_connected_components_min_id_representatives():
    """algo.WCC's LAGr_ConnectedComponents: dense component vector, representative = smallest vertex id of the component, on a
    symmetrised RMAT graph with many isolated vertices, a path, a star and a pair; a directed graph handle is refused"""
    fb.init()
    L = lib()
    from falkordb_b200.grb import Matrix
    A = orc.rmat_csr(13, 2, 21)                       # sparse: hundreds of components
    n = A.nrows
    extra_r = np.array([5000, 5001, 5002, 7000, 7000, 7000, 8100], np.uint64)     # path 5000-5001-5002-5003, star at 7000, pair
    extra_c = np.array([5001, 5002, 5003, 7001, 7002, 7003, 8101], np.uint64)
    r, c, _ = A.tuples()
    rows = np.concatenate([r, c, extra_r, extra_c]).astype(np.uint64)
    cols = np.concatenate([c, r, extra_c, extra_r]).astype(np.uint64)
    S = orc.build_matrix(n, n, rows, cols)
    want = orc.wcc(S)
    m = Matrix.import_csr(n, n, S.p.astype(np.uint64), S.j, None, bool)
    G, h = P(), P(m.h.value)
    m.h = P()
    assert L.LAGraph_New(C.byref(G), C.byref(h), 0, None) == 0           # LAGraph_ADJACENCY_UNDIRECTED
    comp = P()
    assert L.LAGr_ConnectedComponents(C.byref(comp), G, None) == 0
    nv = C.c_uint64(n)
    I, X = np.empty(n, np.uint64), np.empty(n, np.int64)
    check(L.GrB_Vector_extractTuples_INT64(I.ctypes.data, X.ctypes.data, C.byref(nv), comp))     # extract_vector_i64
    assert nv.value == n and np.array_equal(I, np.arange(n)), "the component vector is dense"
    assert np.array_equal(X, want), "representatives differ from the smallest id of each component"
    assert len(np.unique(X)) > 100 and X[5003] == min(X[5000], 5000) and X[7003] == X[7000] and X[8101] == X[8100]
    L.GrB_Vector_free(C.byref(comp))
    L.LAGraph_Delete(C.byref(G), None)
    d = Matrix.import_csr(n, n, A.p.astype(np.uint64), A.j, None, bool)
    G2, h2 = P(), P(d.h.value)
    d.h = P()
    assert L.LAGraph_New(C.byref(G2), C.byref(h2), 1, None) == 0          # directed, symmetry unknown
    assert L.LAGr_ConnectedComponents(C.byref(comp), G2, None) == -1005
    L.LAGraph_Delete(C.byref(G2), None)


@pytest.mark.parametrize("itermax", [1, 3, 10])
def test_lagraph_cdlp_label_propagation(itermax):
    """algo.labelPropagation's LAGraph_cdlp (algo_procedures.rs:1232-1237): synchronous "most frequent neighbour label, smallest on
    ties" from label(v) = v, bit-exact against the oracle after 1, 3 and 10 rounds on a symmetrised RMAT graph (skewed rows, isolated
    vertices, self-edges) joined with the reference flow test's three fully connected triples (tests/flow/test_cdlp.py:83-178)"""
    fb.init()
    L = lib()
    from falkordb_b200.grb import Matrix
    A = orc.rmat_csr(12, 6, 77)
    n = A.nrows
    r, c, _ = A.tuples()
    keep = (r < n - 16) & (c < n - 16)                # the last 16 vertices belong to the triples (and stay isolated beyond them)
    t0 = n - 12
    tr = np.array([t0 + 3 * k + a for k in range(3) for a, b in ((0, 1), (0, 2), (1, 2))], np.uint64)
    tc = np.array([t0 + 3 * k + b for k in range(3) for a, b in ((0, 1), (0, 2), (1, 2))], np.uint64)
    loops = np.array([100, 200, 300], np.uint64)      # self-edges count as neighbours carrying the vertex's own label
    rows = np.concatenate([r[keep], c[keep], tr, tc, loops]).astype(np.uint64)
    cols = np.concatenate([c[keep], r[keep], tc, tr, loops]).astype(np.uint64)
    S = orc.build_matrix(n, n, rows, cols)
    want, rounds = orc.cdlp(S, itermax)
    m = Matrix.import_csr(n, n, S.p.astype(np.uint64), S.j, None, bool)
    G, h = P(), P(m.h.value)
    m.h = P()
    assert L.LAGraph_New(C.byref(G), C.byref(h), 0, None) == 0           # LAGraph_ADJACENCY_UNDIRECTED
    out = P()
    assert L.LAGraph_cdlp(C.byref(out), G, itermax, None) == 0
    nv = C.c_uint64(n)
    I, X = np.empty(n, np.uint64), np.empty(n, np.int64)
    check(L.GrB_Vector_extractTuples_INT64(I.ctypes.data, X.ctypes.data, C.byref(nv), out))      # extract_vector_i64
    assert nv.value == n and np.array_equal(I, np.arange(n)), "the label vector is dense"
    assert np.array_equal(X, want), f"labels differ from the oracle after {rounds} rounds"
    if itermax >= 3:
        for k in range(3):
            assert X[t0 + 3 * k] == X[t0 + 3 * k + 1] == X[t0 + 3 * k + 2] == t0 + 3 * k
    assert X[n - 13] == n - 13, "an isolated vertex keeps its label"
    L.GrB_Vector_free(C.byref(out))
    L.LAGraph_Delete(C.byref(G), None)
    d = Matrix.import_csr(n, n, A.p.astype(np.uint64), A.j, None, bool)
    G2, h2 = P(), P(d.h.value)
    d.h = P()
    assert L.LAGraph_New(C.byref(G2), C.byref(h2), 1, None) == 0          # directed, symmetry unknown
    assert L.LAGraph_cdlp(C.byref(out), G2, itermax, None) == -1005
    L.LAGraph_Delete(C.byref(G2), None)


@pytest.mark.parametrize("nsrc,max_hops,include", [(64, None, False), (200, 2, False), (300, None, True), (5, 1, False), (1, 0, True)])
def test_reach_batch_c_entry(nsrc, max_hops, include):
    """B200_reach_batch (the var-len reach fast path / allShortestPaths BFS phase as one C call) against the level loop on the
    oracle, and entry for entry against the hardware-verified Python composition of the same public calls"""
    from test_gpu_parity import from_dev
    A = orc.rmat_csr(11, 4, 19)
    n = A.nrows
    rng = np.random.default_rng(nsrc)
    src = rng.choice(n, size=nsrc, replace=False)
    dA = to_dev(A)
    R, levels = fb.reach_batch(src, dA, max_hops, include)
    rows = np.arange(nsrc)
    F = orc.build_matrix(nsrc, n, rows, src)
    Ro = orc.build_matrix(nsrc, n, rows, src) if include else orc.build_matrix(nsrc, n, [], [])
    lv = 0
    while max_hops is None or lv < max_hops:
        F = orc.mxm(F, A, Ro, mask_mode=2) if Ro.nnz else orc.mxm(F, A)
        if F.nnz == 0:
            break
        Ro = orc.ewise_add(Ro, F)
        lv += 1
    assert levels == lv
    R.wait()
    assert_same(R, Ro, f"reach_batch nsrc={nsrc} max_hops={max_hops} include={include}")
    R2, levels2 = fb.multi_source_reach(src, dA, max_hops, include)
    R2.wait()
    assert levels2 == levels and from_dev(R2).tuples()[1].tolist() == from_dev(R).tuples()[1].tolist()


def test_repack_output_batches_host_mirror():
    """batch.rs:81, 274-287: <= 1024 rows per output batch, NodeIds + u16 selection vector, order preserved"""
    run_host_test("repack_output_batches")


def test_multiple_edges_flow_test_at_the_tensor_level():
    """tests/flow/test_multiple_edges.py:11-96: edges created and deleted one at a time on one pair -- counts, ids, var-len count"""
    run_host_test("multiple_edges_flow")


def test_var_len_trails_over_a_tensor():
    """cond_var_len_traverse.rs:152-386 with the adjacency fetched from a relationship Tensor through the row iterators"""
    run_host_test("var_len_trails")


def test_tensor_rdb_round_trip_after_device_side_mutations():
    """tensor.rs:1049-1204: batched inserts, a delta fold, bulk deletes with a demotion, then encode -> decode -> rebuild_backward;
    every (src, dst, edge id) is checked against a model kept beside the tensor"""
    run_host_test("tensor_encode_decode_after_mutations")


def test_tensor_bulk_load_host_mirror():
    """GRAPH.BULK into an empty tensor (bulk_insert.rs:497 -> graph.rs:2062 -> tensor.rs:333-447): the device-side build against the
    reference's per-edge insert loop (host mirror), every (src, dst, id), inline values, multi-edge lists, backward matrix"""
    run_host_test("tensor_bulk_load_matches_set_all_from_slices")


@pytest.mark.parametrize("n,count,hot", [(1 << 10, 5000, 64), (1 << 17, 400_000, 1 << 12), (300, 1, 1)])
def test_tensor_bulk_build_against_a_numpy_model(n, count, hot):
    """B200_Tensor_bulk_build at sizes the host mirror's insert loop would take minutes for: forward CSR (columns and values) and the
    multi-edge (key, id) list against numpy on unordered ids"""
    fb.init()
    L = lib()
    from falkordb_b200.grb import Matrix
    rng = np.random.default_rng(count)
    src = rng.integers(0, n, count).astype(np.uint64)
    dst = rng.integers(0, n, count).astype(np.uint64)
    h = rng.random(count) < 0.3                                  # 30 % of the edges fall on few pairs: long id lists
    src[h] = rng.integers(0, hot, int(h.sum())).astype(np.uint64) % np.uint64(n)
    dst[h] = (src[h] * np.uint64(7) + np.uint64(3)) % np.uint64(n)
    ids = rng.permutation(count).astype(np.uint64)               # unordered, id 0 included
    fwd, mk, mi, nm = P(), C.POINTER(U64)(), C.POINTER(U64)(), U64()
    check(L.B200_Tensor_bulk_build(C.byref(fwd), C.byref(mk), C.byref(mi), C.byref(nm), n, n, src.ctypes.data, dst.ctypes.data,
                                   ids.ctypes.data, count))
    M = Matrix(0, 0, np.uint64, _handle=fwd)
    p, j, x = M.export_csr()
    key = (src << np.uint64(32)) | dst
    order = np.lexsort((ids, key))
    ks, iss = key[order], ids[order]
    uk, first, cnt = np.unique(ks, return_index=True, return_counts=True)
    assert M.nvals() == len(uk)
    assert np.array_equal(np.repeat(np.arange(n, dtype=np.uint64), np.diff(p.astype(np.int64))), uk >> np.uint64(32)), "rows"
    assert np.array_equal(j.astype(np.uint64), uk & np.uint64(0xFFFFFFFF)), "columns"
    want_x = np.where(cnt > 1, np.uint64(0xFFFFFFFFFFFFFFFF), iss[first])
    assert np.array_equal(x, want_x), "inline edge id, or MULTI_EDGE for a pair with several edges"
    in_multi = np.repeat(cnt > 1, cnt)
    assert nm.value == int(in_multi.sum())
    if nm.value:
        got_k = np.ctypeslib.as_array(mk, shape=(nm.value,)).copy()
        got_i = np.ctypeslib.as_array(mi, shape=(nm.value,)).copy()
        libc = C.CDLL(None)
        libc.free.argtypes = [C.c_void_p]
        libc.free(mk); libc.free(mi)
        assert np.array_equal(got_k, ks[in_multi]) and np.array_equal(got_i, iss[in_multi]), "the multi-edge list, sorted by (pair, id)"
    else:
        assert not mk and not mi
    bad = dst.copy()
    bad[0] = n
    f2 = P()
    rc = L.B200_Tensor_bulk_build(C.byref(f2), C.byref(mk), C.byref(mi), C.byref(nm), n, n, src.ctypes.data, bad.ctypes.data, ids.ctypes.data, count)
    assert rc != 0 and not f2, "an endpoint outside the matrix is refused and nothing is handed out"
