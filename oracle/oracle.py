"""ctypes loader for the CPU oracle (oracle/grb_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package falkordb_b200 never imports it.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


class _CSR(C.Structure):
    _fields_ = [("nrows", C.c_int64), ("ncols", C.c_int64), ("nnz", C.c_int64),
                ("p", C.POINTER(C.c_int64)), ("j", C.POINTER(C.c_uint32)), ("x", C.POINTER(C.c_uint64))]


def build(force=False):
    so = os.path.join(_HERE, "liborc.so")
    src = os.path.join(_HERE, "grb_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "liborc.so"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(_HERE, "liborc.so")
        if not os.path.exists(so):
            build()
        L = C.CDLL(so)
        P = C.POINTER(_CSR)
        L.orc_build.argtypes = [C.c_int64, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, P]
        L.orc_mask_assign.argtypes = [P, P, P, C.c_int, C.c_int, C.c_int, C.c_int, P]
        L.orc_mxm_anypair.argtypes = [P, P, P, C.c_int, P, C.POINTER(C.c_int64)]
        L.orc_ewise_add.argtypes = [P, P, C.c_int, P]
        L.orc_ewise_mult.argtypes = [P, P, P]
        L.orc_transpose.argtypes = [P, P]
        L.orc_bfs.argtypes = [P, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p]
        L.orc_rmat_edges.argtypes = [C.c_int, C.c_int64, C.c_uint64, C.c_void_p, C.c_void_p]
        L.orc_csr_from_edges.argtypes = [C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, P]
        L.orc_csr_free.argtypes = [P]
        L.orc_chain.argtypes = [P, C.c_void_p, C.c_int64, C.c_int, P, C.POINTER(C.c_int64), C.c_void_p, C.POINTER(C.c_double)]
        L.orc_digest.argtypes = [P, C.c_void_p]
        L.orc_mxv_fp64.argtypes = [P, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_pagerank.argtypes = [P, C.c_void_p, C.c_double, C.c_double, C.c_int, C.c_void_p]
        L.orc_pagerank.restype = C.c_int
        L.orc_last_busy_fraction.restype = C.c_double
        L.orc_last_busy_threads.restype = C.c_int
        L.orc_num_threads.restype = C.c_int
        L.orc_set_num_threads.argtypes = [C.c_int]
        L.orc_parallel_copy.argtypes = [C.c_void_p, C.c_void_p, C.c_int64]
        _LIB = L
    return _LIB


class CSR:
    """Host CSR: p int64[nrows+1], j uint32[nnz] (ascending per row), x uint64[nnz] or None."""

    def __init__(self, nrows, ncols, p, j, x=None):
        self.nrows, self.ncols = int(nrows), int(ncols)
        self.p = np.ascontiguousarray(p, dtype=np.int64)
        self.j = np.ascontiguousarray(j, dtype=np.uint32)
        self.x = None if x is None else np.ascontiguousarray(x, dtype=np.uint64)
        assert self.p.shape[0] == self.nrows + 1

    @property
    def nnz(self):
        return int(self.p[-1])

    def _c(self):
        s = _CSR(self.nrows, self.ncols, self.nnz,
                 self.p.ctypes.data_as(C.POINTER(C.c_int64)),
                 self.j.ctypes.data_as(C.POINTER(C.c_uint32)),
                 self.x.ctypes.data_as(C.POINTER(C.c_uint64)) if self.x is not None else None)
        return s

    def tuples(self):
        rows = np.repeat(np.arange(self.nrows, dtype=np.uint64), np.diff(self.p))
        return rows, self.j.astype(np.uint64), self.x

    def tuple_set(self):
        r, c, x = self.tuples()
        if x is None:
            return set(zip(r.tolist(), c.tolist()))
        return set(zip(r.tolist(), c.tolist(), x.tolist()))

    def to_scipy(self):
        import scipy.sparse as sp
        data = np.ones(self.nnz, dtype=np.int64) if self.x is None else self.x.astype(np.int64)
        return sp.csr_matrix((data, self.j.astype(np.int64), self.p), shape=(self.nrows, self.ncols))

    @staticmethod
    def from_scipy(m, values=False):
        m = m.tocsr()
        m.sort_indices()
        return CSR(m.shape[0], m.shape[1], m.indptr.astype(np.int64), m.indices.astype(np.uint32),
                   m.data.astype(np.uint64) if values else None)

    @staticmethod
    def empty(nrows, ncols, values=False):
        return CSR(nrows, ncols, np.zeros(nrows + 1, np.int64), np.zeros(0, np.uint32),
                   np.zeros(0, np.uint64) if values else None)

    def __eq__(self, o):
        return (self.nrows == o.nrows and self.ncols == o.ncols and np.array_equal(self.p, o.p)
                and np.array_equal(self.j, o.j)
                and ((self.x is None) == (o.x is None))
                and (self.x is None or np.array_equal(self.x, o.x)))


def _take(out):
    """Copy a C-allocated orc_csr into numpy-owned arrays and free the C side."""
    n, nnz = out.nrows, out.nnz
    p = np.ctypeslib.as_array(out.p, shape=(n + 1,)).copy()
    j = np.ctypeslib.as_array(out.j, shape=(nnz,)).copy() if nnz else np.zeros(0, np.uint32)
    x = None
    if bool(out.x):
        x = np.ctypeslib.as_array(out.x, shape=(nnz,)).copy() if nnz else np.zeros(0, np.uint64)
    res = CSR(n, out.ncols, p, j, x)
    lib().orc_csr_free(C.byref(out))
    return res


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def build_matrix(nrows, ncols, I, J, X=None):
    I = np.ascontiguousarray(I, dtype=np.uint64)
    J = np.ascontiguousarray(J, dtype=np.uint64)
    X = None if X is None else np.ascontiguousarray(X, dtype=np.uint64)
    out = _CSR()
    rc = lib().orc_build(nrows, ncols, len(I), _ptr(I), _ptr(J), _ptr(X), C.byref(out))
    if rc:
        raise IndexError("GrB_INDEX_OUT_OF_BOUNDS")
    return _take(out)


def mask_assign(Cold, T, M=None, comp=False, structural=False, replace=False, accum=False):
    out = _CSR()
    c = Cold._c() if Cold is not None else None
    t = T._c()
    m = M._c() if M is not None else None
    lib().orc_mask_assign(C.byref(c) if c else None, C.byref(t), C.byref(m) if m else None,
                          int(comp), int(structural), int(replace), int(accum), C.byref(out))
    return _take(out)


def mxm(A, B, M=None, mask_mode=0, return_flops=False):
    """T = pattern(A*B) over ANY_PAIR; mask_mode 0 none / 1 keep M / 2 drop M (structural)."""
    out = _CSR()
    a, b = A._c(), B._c()
    m = M._c() if M is not None else None
    fl = C.c_int64(0)
    rc = lib().orc_mxm_anypair(C.byref(a), C.byref(b), C.byref(m) if m else None, mask_mode,
                               C.byref(out), C.byref(fl))
    if rc:
        raise ValueError("GrB_DIMENSION_MISMATCH")
    r = _take(out)
    return (r, fl.value) if return_flops else r


def chain(A, sources, hops, keep=True):
    """CondTraverse's F(i, src_i) = 1; F <- F*A `hops` times (cond_traverse.rs:600-608), entirely in C.
    Returns (result CSR or None, flops, digest uint64[3], busy fraction of the host threads)."""
    src = np.ascontiguousarray(sources, dtype=np.uint64)
    out = _CSR()
    a = A._c()
    fl = C.c_int64(0)
    dg = np.zeros(3, np.uint64)
    busy = C.c_double(0.0)
    rc = lib().orc_chain(C.byref(a), _ptr(src), len(src), hops, C.byref(out) if keep else None, C.byref(fl), _ptr(dg),
                         C.byref(busy))
    if rc:
        raise IndexError("GrB_INVALID_INDEX")
    return (_take(out) if keep else None), fl.value, dg, busy.value


def digest(A):
    """(nvals, sum mix(row << 32 | col), sum mix(key + GOLD * (position + 1))): order-sensitive digest of a CSR pattern;
    falkordb_b200's B200_Matrix_digest computes the same three numbers on the device."""
    dg = np.zeros(3, np.uint64)
    a = A._c()
    lib().orc_digest(C.byref(a), _ptr(dg))
    return dg


def busy_fraction():
    return lib().orc_last_busy_fraction()


def ewise_add(A, B, keep_values=False):
    out = _CSR()
    a, b = A._c(), B._c()
    rc = lib().orc_ewise_add(C.byref(a), C.byref(b), int(keep_values), C.byref(out))
    if rc:
        raise ValueError("GrB_DIMENSION_MISMATCH")
    return _take(out)


def ewise_mult(A, B):
    out = _CSR()
    a, b = A._c(), B._c()
    rc = lib().orc_ewise_mult(C.byref(a), C.byref(b), C.byref(out))
    if rc:
        raise ValueError("GrB_DIMENSION_MISMATCH")
    return _take(out)


def transpose(A):
    out = _CSR()
    a = A._c()
    lib().orc_transpose(C.byref(a), C.byref(out))
    return _take(out)


def bfs(A, src, max_level=-1, want_parent=True):
    n = A.nrows
    level = np.empty(n, np.int64)
    parent = np.empty(n, np.int64) if want_parent else None
    a = A._c()
    rc = lib().orc_bfs(C.byref(a), src, max_level, _ptr(level), _ptr(parent))
    if rc:
        raise IndexError("GrB_INVALID_INDEX")
    return level, parent


def rmat_edges(scale, nedges, seed):
    I = np.empty(nedges, np.uint64)
    J = np.empty(nedges, np.uint64)
    lib().orc_rmat_edges(scale, nedges, seed, _ptr(I), _ptr(J))
    return I, J


def rmat_csr(scale, edge_factor=16, seed=1):
    """Deduplicated, self-loop-free directed RMAT pattern CSR."""
    n = 1 << scale
    I, J = rmat_edges(scale, n * edge_factor, seed)
    out = _CSR()
    lib().orc_csr_from_edges(n, len(I), _ptr(I), _ptr(J), C.byref(out))
    return _take(out)


def mxv_fp64(A, x, use_values=True, present=None):
    """y = A*x over PLUS_TIMES (A.x = IEEE-754 bit patterns of doubles) or PLUS_SECOND (use_values=False).  Returns (y, ypresent)."""
    x = np.ascontiguousarray(x, dtype=np.float64)
    pr = None if present is None else np.ascontiguousarray(present, dtype=np.uint8)
    y = np.zeros(A.nrows, np.float64)
    yp = np.zeros(A.nrows, np.uint8)
    a = A._c()
    lib().orc_mxv_fp64(C.byref(a), int(use_values), _ptr(x), _ptr(pr), _ptr(y), _ptr(yp))
    return y, yp.astype(bool)


def pagerank(A, damping=0.85, tol=1e-4, itermax=100):
    """LAGr_PageRank on the pattern of A (directed; rank flows along out-edges).  Returns (scores float64[n], iterations)."""
    AT = transpose(pattern(A))
    deg = np.ascontiguousarray(np.diff(A.p), dtype=np.int64)
    r = np.zeros(A.nrows, np.float64)
    at = AT._c()
    it = lib().orc_pagerank(C.byref(at), _ptr(deg), damping, tol, itermax, _ptr(r))
    return r, it


def wcc(A):
    """LAGr_ConnectedComponents as algo.WCC consumes it (algo_procedures.rs:838-846): component(i) = representative of i's
    component; FastSV's min-hooking makes that the SMALLEST vertex id of the component.  A: symmetric pattern.  Restated with
    scipy's connected_components (an independent implementation) + a min per label."""
    import scipy.sparse.csgraph as cg
    ncomp, lab = cg.connected_components(A.to_scipy(), directed=False)
    rep = np.full(ncomp, A.nrows, np.int64)
    np.minimum.at(rep, lab, np.arange(A.nrows, dtype=np.int64))
    return rep[lab]


def cdlp(A, itermax=10):
    """LAGraph_cdlp as algo.labelPropagation consumes it (algo_procedures.rs:1232-1237; LAGraph experimental/algorithm/LAGraph_cdlp.c,
    not vendored; the algorithm is LDBC Graphalytics' CDLP): label(v) = v; each round every vertex takes the most frequent label among
    the entries of its row, the smallest such label on ties, synchronously; a vertex with an empty row keeps its label; at most
    `itermax` rounds, stopping early at a fixed point.  A: the symmetric pattern the reference builds.  Returns (labels, rounds)."""
    n = A.nrows
    L = np.arange(n, dtype=np.int64)
    rows = np.repeat(np.arange(n, dtype=np.int64), np.diff(A.p))
    rounds = 0
    if A.nnz == 0:
        return L, 0
    while rounds < itermax:
        key, cnt = np.unique(rows * (n + 1) + L[A.j], return_counts=True)
        r, lab = key // (n + 1), key % (n + 1)
        order = np.lexsort((lab, -cnt, r))                  # by row, then count descending, then label ascending
        first = np.ones(len(order), bool)
        first[1:] = r[order][1:] != r[order][:-1]
        new = L.copy()
        new[r[order][first]] = lab[order][first]
        rounds += 1
        same = np.array_equal(new, L)
        L = new
        if same:
            break
    return L, rounds


def num_threads():
    return lib().orc_num_threads()


def spread(A):
    """A copy of A whose arrays were first touched by every thread of the team (2 MB pieces, round-robin): on a multi-socket host
    the graph is then interleaved over the memory controllers rather than resident on the node of the thread that built it."""
    def cp(a):
        if a is None:
            return None
        out = np.empty_like(a)                       # fresh, untouched pages
        if a.nbytes:
            lib().orc_parallel_copy(out.ctypes.data, a.ctypes.data, a.nbytes)
        return out
    return CSR(A.nrows, A.ncols, cp(A.p), cp(A.j), cp(A.x))


# ---- wrapper-level algebra the reference documents in-tree, composed from the primitives ----

def delta_lmxm(F, m, dp, dm):
    """Matrix::<bool>::delta_lmxm, matrix.rs:1317-1402: F <- (F*(m u dp)) <!(F*dm)>."""
    if dp.nnz == 0 and dm.nnz == 0:
        return mxm(F, m)
    mask = None
    if dm.nnz > 0:
        mk = mxm(F, dm)
        if mk.nnz > 0:
            mask = mk
    accum = None
    if dp.nnz > 0:
        ac = mxm(F, dp)
        if ac.nnz > 0:
            accum = ac
    out = mxm(F, m, mask, 2) if mask is not None else mxm(F, m)
    if accum is not None:
        out = ewise_add(out, accum)
    return out


def pattern(A):
    return CSR(A.nrows, A.ncols, A.p, A.j, None)
