"""CondTraverse's batched F*A path through the C++ host mirror (falkordb_b200/csrc/host/cond_traverse.hpp,
mirroring graph/src/runtime/ops/cond_traverse.rs:452-751): the README MotoGP example (BASELINE config 1) and a
3-hop chain with label filters and a dirty snapshot, checked against the oracle."""
import ctypes as C
import os

import numpy as np
import pytest

import oracle as orc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def host():
    L = C.CDLL(os.path.join(ROOT, "falkordb_b200", "libfdbhost.so"))
    L.fdbh_run_test.argtypes = [C.c_char_p]
    L.fdbh_last_message.restype = C.c_char_p
    L.fdbh_vm_from_csr.argtypes = [C.c_uint64, C.c_uint64, C.c_void_p, C.c_void_p]
    L.fdbh_vm_from_csr.restype = C.c_void_p
    L.fdbh_vm_new.argtypes = [C.c_uint64, C.c_uint64]
    L.fdbh_vm_new.restype = C.c_void_p
    L.fdbh_vm_free.argtypes = [C.c_void_p]
    L.fdbh_vm_set.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64]
    L.fdbh_vm_remove.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64]
    L.fdbh_vm_nvals.argtypes = [C.c_void_p]
    L.fdbh_vm_nvals.restype = C.c_int64
    L.fdbh_expand_batch.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64, C.c_void_p,
                                    C.c_uint64, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(C.c_uint64)]
    L.fdbh_expand_batch_fused.argtypes = L.fdbh_expand_batch.argtypes
    L.fdbh_free.argtypes = [C.c_void_p]
    return L


def test_motogp_readme_example():
    """README.md:85-110: Yamaha's rider is Valentino Rossi, count 1; plus a second hop and a pending delete"""
    L = host()
    assert L.fdbh_run_test(b"motogp_two_hop") == 0, L.fdbh_last_message().decode()


def vm_from(L, c):
    p = c.p.astype(np.uint64)
    h = L.fdbh_vm_from_csr(c.nrows, c.ncols, p.ctypes.data, c.j.ctypes.data if c.nnz else None)
    assert h, L.fdbh_last_message().decode()
    return h


def expand(L, src, hops, sl=(), dl=(), fused=False):
    src = np.ascontiguousarray(src, dtype=np.uint64)
    arr = lambda hs: (C.c_void_p * max(1, len(hs)))(*hs)
    rows, dest, n = C.c_void_p(), C.c_void_p(), C.c_uint64()
    rc = (L.fdbh_expand_batch_fused if fused else L.fdbh_expand_batch)(src.ctypes.data, len(src), arr(hops), len(hops), arr(sl), len(sl), arr(dl), len(dl),
                             C.byref(rows), C.byref(dest), C.byref(n))
    assert rc == 0, L.fdbh_last_message().decode()
    r = np.ctypeslib.as_array(C.cast(rows, C.POINTER(C.c_uint64)), shape=(n.value,)).copy() if n.value else np.zeros(0, np.uint64)
    d = np.ctypeslib.as_array(C.cast(dest, C.POINTER(C.c_uint64)), shape=(n.value,)).copy() if n.value else np.zeros(0, np.uint64)
    L.fdbh_free(rows); L.fdbh_free(dest)
    return r, d


def test_expand_batch_three_hops_with_labels_and_dirty_snapshot():
    L = host()
    A = orc.rmat_csr(11, 8, 9)
    n = A.nrows
    rng = np.random.default_rng(2)
    src = rng.choice(np.nonzero(np.diff(A.p))[0], 200, replace=False)
    vA = vm_from(L, A)
    # clean snapshot, 3 fused hops (fuse_anonymous_traverse: chain of storage-direction hops)
    r, d = expand(L, src, [vA, vA, vA])
    F = orc.build_matrix(len(src), n, np.arange(len(src)), src)
    W = F
    for _ in range(3):
        W = orc.mxm(W, A)
    wr, wc, _ = W.tuples()
    assert np.array_equal(r, wr) and np.array_equal(d, wc)          # (row_i, dest) ascending, as the op relies on
    # labels: source pre-filter and destination post-filter (diagonal label matrices, graph.rs:1191)
    lab_s = np.sort(rng.choice(n, n // 2, replace=False))
    lab_d = np.sort(rng.choice(n, n // 3, replace=False))
    vs = vm_from(L, orc.build_matrix(n, n, lab_s, lab_s))
    vd = vm_from(L, orc.build_matrix(n, n, lab_d, lab_d))
    r, d = expand(L, src, [vA, vA], [vs], [vd])
    keep = np.isin(src, lab_s)
    F2 = orc.build_matrix(len(src), n, np.arange(len(src))[keep], src[keep])
    W2 = orc.mxm(orc.mxm(F2, A), A)
    wr, wc, _ = W2.tuples()
    sel = np.isin(wc, lab_d)
    assert np.array_equal(r, wr[sel]) and np.array_equal(d, wc[sel])
    # dirty snapshot: pending adds and deletes go through delta_lmxm's 3-mxm form (matrix.rs:1342-1400)
    vB = vm_from(L, A)
    rows, cols, _ = A.tuples()
    dels = rng.choice(A.nnz, 300, replace=False)
    for q in dels:
        assert L.fdbh_vm_remove(vB, int(rows[q]), int(cols[q])) == 0
    adds = [(int(a), int(b)) for a, b in zip(rng.integers(0, n, 300), rng.integers(0, n, 300))]
    have = A.tuple_set()
    adds = sorted({t for t in adds if t not in have})
    for a, b in adds:
        assert L.fdbh_vm_set(vB, a, b) == 0
    dm = orc.build_matrix(n, n, rows[dels], cols[dels])
    dp = orc.build_matrix(n, n, [a for a, _ in adds], [b for _, b in adds])
    assert L.fdbh_vm_nvals(vB) == A.nnz - len(dels) + dp.nnz
    r, d = expand(L, src, [vB])
    W3 = orc.delta_lmxm(F, A, dp, dm)
    wr, wc, _ = W3.tuples()
    assert np.array_equal(r, wr) and np.array_equal(d, wc)
    for h in (vA, vB, vs, vd):
        L.fdbh_vm_free(h)


def test_destination_label_filter_fused_as_a_diagonal_hop():
    """F*A*L_dst with the label filter on the device (one more delta_lmxm against the diagonal label
    VersionedMatrix, clean and dirty) gives exactly the pairs the per-output node_has_label probes keep"""
    L = host()
    A = orc.rmat_csr(11, 8, 13)
    n = A.nrows
    rng = np.random.default_rng(4)
    src = rng.choice(np.nonzero(np.diff(A.p))[0], 300, replace=False)
    vA = vm_from(L, A)
    lab = np.sort(rng.choice(n, n // 3, replace=False))
    vd = vm_from(L, orc.build_matrix(n, n, lab, lab))
    F = orc.build_matrix(len(src), n, np.arange(len(src)), src)
    W = orc.mxm(orc.mxm(F, A), A)
    wr, wc, _ = W.tuples()
    for fused in (False, True):
        r, d = expand(L, src, [vA, vA], [], [vd], fused=fused)
        sel = np.isin(wc, lab)
        assert np.array_equal(r, wr[sel]) and np.array_equal(d, wc[sel]), f"fused={fused}"
    # dirty label matrix: some labels dropped, some added since the base was written
    drop = rng.choice(lab, 100, replace=False)
    add = np.setdiff1d(rng.choice(n, 200, replace=False), lab)
    for v in drop:
        assert L.fdbh_vm_remove(vd, int(v), int(v)) == 0
    for v in add:
        assert L.fdbh_vm_set(vd, int(v), int(v)) == 0
    lab2 = np.union1d(np.setdiff1d(lab, drop), add)
    for fused in (False, True):
        r, d = expand(L, src, [vA, vA], [], [vd], fused=fused)
        sel = np.isin(wc, lab2)
        assert np.array_equal(r, wr[sel]) and np.array_equal(d, wc[sel]), f"dirty labels, fused={fused}"
    for h in (vA, vd):
        L.fdbh_vm_free(h)
