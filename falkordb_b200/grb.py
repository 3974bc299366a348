"""Host-side mirror of the reference's GraphBLAS wrapper, method for method:
``Matrix<T>`` in graph/src/graph/graphblas/matrix.rs (new :1119/:1214, build :1186/:1281, set/get
:1143-1172/:1248-1275, lmxm/rmxm :930-968, delta_lmxm :1317-1402, element_wise_add :852, element_wise_multiply
:876, set_pattern :906, transpose :633, remove_all/select :824-845, wait :781, dup :1062, grown :700, iter :1471).
Everything is a direct call into libb200grb.so through the GraphBLAS C ABI (include/b200grb.h)."""
import ctypes as C
import enum

import numpy as np

from ._lib import lib, obj, check, GrbError, P, U64, I64

GxB_SPARSITY_STATUS, GxB_SPARSITY_CONTROL, GxB_HYPER_HASH, GxB_WILL_WAIT = 7034, 7036, 7048, 7076
GrB_STORAGE_ORIENTATION_HINT = 100
GxB_HYPERSPARSE, GxB_SPARSE = 1, 2
GrB_MATERIALIZE = 1
U64_MAX = (1 << 64) - 1


class Descriptor(enum.Enum):
    """matrix.rs:223-255"""
    T0 = "T0"; T1 = "T1"; T0T1 = "T0T1"; C = "C"; CT0 = "CT0"; CT1 = "CT1"; CT0T1 = "CT0T1"
    S = "S"; ST0 = "ST0"; ST1 = "ST1"; ST0T1 = "ST0T1"; SC = "SC"; SCT0 = "SCT0"; SCT1 = "SCT1"; SCT0T1 = "SCT0T1"
    R = "R"; RT0 = "RT0"; RT1 = "RT1"; RT0T1 = "RT0T1"; RC = "RC"; RCT0 = "RCT0"; RCT1 = "RCT1"; RCT0T1 = "RCT0T1"
    RS = "RS"; RST0 = "RST0"; RST1 = "RST1"; RST0T1 = "RST0T1"; RSC = "RSC"; RSCT0 = "RSCT0"; RSCT1 = "RSCT1"
    RSCT0T1 = "RSCT0T1"


def _desc(d):
    if d is None:
        return None
    return obj("GrB_DESC_" + d.value)


def init():
    """matrix.rs:116-185 (GxB_init NONBLOCKING + JIT control + LAGraph_Init)"""
    L = lib()
    check(L.GxB_init(0, None, None, None, None))
    check(L.GrB_Global_set_INT32(obj("GrB_GLOBAL"), 2, 7029))
    assert L.LAGraph_Init(None) == 0


def _u64arr(a):
    return np.ascontiguousarray(a, dtype=np.uint64)


class Matrix:
    """``Matrix<bool>`` (dtype=bool) or ``Matrix<u64>`` (dtype='u64')."""

    def __init__(self, nrows, ncols, dtype=bool, _handle=None):
        self.dtype = bool if dtype in (bool, "bool") else "u64"
        if _handle is not None:
            self.h = _handle
            return
        h = P()
        check(lib().GrB_Matrix_new(C.byref(h), obj("GrB_BOOL" if self.dtype is bool else "GrB_UINT64"), nrows, ncols))
        self.h = h
        # pin_sparse, matrix.rs:405-426
        check(lib().GrB_Matrix_set_INT32(self.h, GxB_SPARSE | GxB_HYPERSPARSE, GxB_SPARSITY_CONTROL))
        check(lib().GrB_Matrix_set_INT32(self.h, 0, GrB_STORAGE_ORIENTATION_HINT))

    def __del__(self):
        try:
            if getattr(self, "h", None) is not None and self.h.value:
                lib().GrB_Matrix_free(C.byref(self.h))
        except Exception:
            pass

    # ---- constructors -------------------------------------------------------------------------
    @staticmethod
    def import_csr(nrows, ncols, p, j, x=None, dtype=bool):
        p = np.ascontiguousarray(p, dtype=np.uint64)
        j = np.ascontiguousarray(j, dtype=np.uint32)
        xx = None if x is None else np.ascontiguousarray(x, dtype=np.uint64)
        h = P()
        dt = bool if dtype in (bool, "bool") else "u64"
        check(lib().B200_Matrix_import_CSR(C.byref(h), obj("GrB_BOOL" if dt is bool else "GrB_UINT64"), nrows, ncols,
                                           p.ctypes.data, j.ctypes.data if len(j) else None,
                                           None if xx is None else xx.ctypes.data, 0))
        return Matrix(nrows, ncols, dt, _handle=h)

    def export_csr(self):
        n, nnz = self.nrows(), self.nvals()
        p = np.empty(n + 1, np.uint64)
        j = np.empty(nnz, np.uint32)
        x = np.empty(nnz, np.uint64) if self.dtype != bool else None
        check(lib().B200_Matrix_export_CSR(self.h, p.ctypes.data, j.ctypes.data if nnz else None,
                                           None if x is None else x.ctypes.data, 0))
        return p.astype(np.int64), j, x

    def export_bitmap(self, out=None):
        """Row-major packed bitmap (B200_Matrix_export_bitmap): uint64[nrows, ceil(ncols/64)], bit (j & 63) of word j >> 6 of
        row i set iff (i, j) is an entry.  Returns (bitmap, nvals).  `out` may be a caller (pinned) buffer of that shape."""
        nr, nc = self.nrows(), self.ncols()
        wpr = (nc + 63) // 64
        if out is None:
            out = np.empty((nr, wpr), np.uint64)
        assert out.dtype == np.uint64 and out.size >= nr * wpr and out.flags.c_contiguous
        nv = U64()
        check(lib().B200_Matrix_export_bitmap(self.h, out.ctypes.data, wpr, C.byref(nv), 0))
        return out, nv.value

    def export_bitmap_async(self, out):
        """Non-blocking export_bitmap into `out` (pinned uint64[nrows, ceil(ncols/64)]): returns a ticket for `wait_ticket`;
        the copy runs on the library's second stream, so further mxm calls overlap it.  The matrix may be dropped at once."""
        nr, nc = self.nrows(), self.ncols()
        wpr = (nc + 63) // 64
        assert out.dtype == np.uint64 and out.size >= nr * wpr and out.flags.c_contiguous
        t = P()
        check(lib().B200_Matrix_export_bitmap_async(self.h, out.ctypes.data, wpr, C.byref(t)))
        return t

    def export_auto(self, out_bitmap=None):
        """The result hand-off a traversal operator makes: bitmap when the result is denser than one entry per 32 slots
        (1 bit per slot beats a 4-byte column index per entry), CSR otherwise -- the rule SuiteSparse applies when it
        switches a matrix to GxB_BITMAP.  Returns ("bitmap", bitmap, nvals) or ("csr", (p, j, x), nvals)."""
        nv = self.nvals()
        if nv * 32 > self.nrows() * self.ncols():
            bm, nv2 = self.export_bitmap(out_bitmap)
            return "bitmap", bm, nv2
        return "csr", self.export_csr(), nv

    def into_hyper(self):
        """matrix.rs:558-575"""
        check(lib().GrB_Matrix_set_INT32(self.h, GxB_HYPERSPARSE, GxB_SPARSITY_CONTROL))
        check(lib().GrB_Matrix_set_INT32(self.h, 0, GxB_HYPER_HASH))
        return self

    # ---- shape / status -----------------------------------------------------------------------
    def nrows(self):
        v = U64()
        check(lib().GrB_Matrix_nrows(C.byref(v), self.h))
        return v.value

    def ncols(self):
        v = U64()
        check(lib().GrB_Matrix_ncols(C.byref(v), self.h))
        return v.value

    def nvals(self):
        v = U64()
        check(lib().GrB_Matrix_nvals(C.byref(v), self.h))
        return v.value

    def pending(self):
        v = C.c_int32()
        check(lib().GrB_Matrix_get_INT32(self.h, C.byref(v), GxB_WILL_WAIT))
        return v.value == 1

    def sparsity_status(self):
        v = C.c_int32()
        check(lib().GrB_Matrix_get_INT32(self.h, C.byref(v), GxB_SPARSITY_STATUS))
        return {1: "hypersparse", 2: "sparse", 4: "bitmap", 8: "full"}.get(v.value, "unknown")

    def is_iso(self):
        v = C.c_bool()
        check(lib().GxB_Matrix_iso(C.byref(v), self.h))
        return v.value

    def memory_usage(self):
        v = C.c_size_t()
        check(lib().GxB_Matrix_memoryUsage(C.byref(v), self.h))
        return v.value

    def wait(self):
        check(lib().GrB_Matrix_wait(self.h, GrB_MATERIALIZE))

    def clear(self):
        check(lib().GrB_Matrix_clear(self.h))

    def resize(self, nrows, ncols):
        check(lib().GrB_Matrix_resize(self.h, nrows, ncols))

    def dup(self):
        h = P()
        check(lib().GrB_Matrix_dup(C.byref(h), self.h))
        return Matrix(0, 0, self.dtype, _handle=h)

    def grown(self, nrows, ncols):
        r0, c0 = self.nrows(), self.ncols()
        assert nrows >= r0 and ncols >= c0, f"grown must not shrink: {r0}x{c0} -> {nrows}x{ncols}"
        out = self.dup()
        if nrows != r0 or ncols != c0:
            out.resize(nrows, ncols)
        return out

    # ---- element access -----------------------------------------------------------------------
    def set(self, i, j, value=True):
        if self.dtype is bool:
            check(lib().GrB_Matrix_setElement_BOOL(self.h, bool(value), i, j))
        else:
            check(lib().GrB_Matrix_setElement_UINT64(self.h, int(value), i, j))

    def get(self, i, j):
        if self.dtype is bool:
            v = C.c_bool()
            info = lib().GrB_Matrix_extractElement_BOOL(C.byref(v), self.h, i, j)
        else:
            v = U64()
            info = lib().GrB_Matrix_extractElement_UINT64(C.byref(v), self.h, i, j)
        if info == 0:
            return v.value
        if info == 1:
            return None
        check(info)

    def contains(self, i, j):
        return lib().GxB_Matrix_isStoredElement(self.h, i, j) == 0

    def remove(self, i, j):
        check(lib().GrB_Matrix_removeElement(self.h, i, j))

    def build(self, rows, cols, vals=None):
        rows, cols = _u64arr(rows), _u64arr(cols)
        assert len(rows) == len(cols)
        if len(rows) == 0:
            return
        if self.dtype is bool:
            s = P()
            check(lib().GrB_Scalar_new(C.byref(s), obj("GrB_BOOL")))
            check(lib().GrB_Scalar_setElement_BOOL(s, True))
            try:
                check(lib().GxB_Matrix_build_Scalar(self.h, rows.ctypes.data, cols.ctypes.data, s, len(rows)))
            finally:
                lib().GrB_Scalar_free(C.byref(s))
        else:
            vals = _u64arr(vals)
            check(lib().GrB_Matrix_build_UINT64(self.h, rows.ctypes.data, cols.ctypes.data, vals.ctypes.data, len(rows),
                                                obj("GxB_ANY_UINT64")))

    # ---- bulk algebra -------------------------------------------------------------------------
    def lmxm(self, b):
        check(lib().GrB_mxm(self.h, None, None, obj("GxB_ANY_PAIR_BOOL"), self.h, b.h, None))

    def rmxm(self, b):
        check(lib().GrB_mxm(self.h, None, None, obj("GxB_ANY_PAIR_BOOL"), b.h, self.h, None))

    def mxm(self, a, b, mask=None, descriptor=None):
        check(lib().GrB_mxm(self.h, mask.h if mask is not None else None, None, obj("GxB_ANY_PAIR_BOOL"), a.h, b.h,
                            _desc(descriptor)))

    def delta_lmxm(self, m, dp, dm):
        """matrix.rs:1317-1402, statement for statement."""
        dp.wait()
        dm.wait()
        dp_nvals, dm_nvals = dp.nvals(), dm.nvals()
        if dp_nvals == 0 and dm_nvals == 0:
            self.lmxm(m)
            return
        nrows, ncols = self.nrows(), m.ncols()
        mask = None
        if dm_nvals > 0:
            mk = Matrix(nrows, ncols, bool)
            mk.mxm(self, dm)
            if mk.nvals() > 0:
                mask = mk
        accum = None
        if dp_nvals > 0:
            ac = Matrix(nrows, ncols, bool)
            ac.mxm(self, dp)
            if ac.nvals() > 0:
                accum = ac
        if mask is not None:
            self.mxm(self, m, mask, Descriptor.RSC)
        else:
            self.mxm(self, m)
        if accum is not None:
            self.element_wise_add(None, None, accum, None)

    def element_wise_add(self, mask=None, a=None, b=None, descriptor=None):
        op = obj("GxB_ANY_BOOL") if self.dtype is bool else obj("GrB_SECOND_UINT64")
        check(lib().GrB_Matrix_eWiseAdd_BinaryOp(self.h, mask.h if mask is not None else None, None, op,
                                                 (a or self).h, (b or self).h, _desc(descriptor)))

    def element_wise_multiply(self, mask=None, a=None, b=None, descriptor=None):
        check(lib().GrB_Matrix_eWiseMult_Semiring(self.h, mask.h if mask is not None else None, None,
                                                  obj("GxB_ANY_PAIR_BOOL"), (a or self).h, (b or self).h, _desc(descriptor)))

    def intersection_nvals(self, b):
        t = Matrix(self.nrows(), self.ncols(), bool)
        check(lib().GrB_Matrix_eWiseMult_Semiring(t.h, None, None, obj("GxB_ANY_PAIR_BOOL"), self.h, b.h, None))
        return t.nvals()

    def set_pattern(self, mask, a, descriptor=None):
        check(lib().GrB_Matrix_apply(self.h, mask.h if mask is not None else None, obj("GxB_ANY_BOOL"), obj("GxB_ONE_BOOL"),
                                     a.h, _desc(descriptor)))

    def transpose(self):
        t = Matrix(self.ncols(), self.nrows(), self.dtype)
        check(lib().GrB_transpose(t.h, None, None, self.h, None))
        return t

    def remove_all(self, b):
        check(lib().GrB_transpose(self.h, b.h, None, self.h, obj("GrB_DESC_RCT0")))

    def select(self, mask, a):
        check(lib().GrB_transpose(self.h, mask.h, None, a.h, obj("GrB_DESC_RCT0")))

    # ---- iteration (matrix.rs:1471-1605: the reference's loop over the C iterator, verbatim) ----
    def iter(self, min_row=0, max_row=U64_MAX):
        L = lib()
        it = P()
        check(L.GxB_Iterator_new(C.byref(it)))
        try:
            check(L.GxB_rowIterator_attach(it, self.h, None))
            info = L.GxB_rowIterator_seekRow(it, min_row)
            while info == 1 and L.GxB_rowIterator_getRowIndex(it) < max_row:
                info = L.GxB_rowIterator_nextRow(it)
            depleted = info != 0 or L.GxB_rowIterator_getRowIndex(it) > max_row
            while not depleted:
                row, col = L.GxB_rowIterator_getRowIndex(it), L.GxB_rowIterator_getColIndex(it)
                item = (row, col) if self.dtype is bool else (row, col, L.GxB_Iterator_get_UINT64(it))
                if L.GxB_rowIterator_nextCol(it) != 0:
                    info = L.GxB_rowIterator_nextRow(it)
                    while info == 1 and L.GxB_rowIterator_getRowIndex(it) < max_row:
                        info = L.GxB_rowIterator_nextRow(it)
                    depleted = info != 0 or L.GxB_rowIterator_getRowIndex(it) > max_row
                yield item
        finally:
            L.GxB_Iterator_free(C.byref(it))

    def hyper_vector_count(self):
        if self.sparsity_status() != "hypersparse":
            return None
        L = lib()
        it = P()
        check(L.GxB_Iterator_new(C.byref(it)))
        check(L.GxB_rowIterator_attach(it, self.h, None))
        k = L.GxB_rowIterator_kount(it)
        L.GxB_Iterator_free(C.byref(it))
        return k

    def tuple_set(self):
        return set(self.iter())

    def extract_tuples(self):
        n = self.nvals()
        I, J = np.empty(n, np.uint64), np.empty(n, np.uint64)
        nv = U64(n)
        if self.dtype is bool:
            check(lib().GrB_Matrix_extractTuples_BOOL(I.ctypes.data, J.ctypes.data, None, C.byref(nv), self.h))
            return I, J, None
        X = np.empty(n, np.uint64)
        check(lib().GrB_Matrix_extractTuples_UINT64(I.ctypes.data, J.ctypes.data, X.ctypes.data, C.byref(nv), self.h))
        return I, J, X

    def extract_pairs(self, rows, cols):
        """Batched `get`: (found bool[n], values u64[n]) for the pairs (rows[t], cols[t]) -- ExpandInto's probe."""
        rows, cols = _u64arr(rows), _u64arr(cols)
        found = np.zeros(len(rows), np.uint8)
        vals = np.zeros(len(rows), np.uint64)
        check(lib().B200_Matrix_extract_pairs(self.h, rows.ctypes.data, cols.ctypes.data, len(rows), found.ctypes.data,
                                              vals.ctypes.data))
        return found.astype(bool), vals

    def digest(self):
        """(nvals, sum mix(key), sum mix(key + GOLD * position)) of the pattern in CSR order -- oracle.digest's twin."""
        d = np.zeros(3, np.uint64)
        check(lib().B200_Matrix_digest(self.h, d.ctypes.data))
        return d

    def prepare(self, want_transpose=True):
        check(lib().B200_Matrix_prepare(self.h, int(want_transpose)))
        return self


def rmat(scale, edge_factor=16, seed=1):
    h = P()
    check(lib().B200_Matrix_rmat(C.byref(h), scale, edge_factor, seed))
    return Matrix(0, 0, bool, _handle=h)


def bfs(A, src, max_level=-1, want_parent=True):
    n = A.nrows()
    level = np.empty(n, np.int64)
    parent = np.empty(n, np.int64) if want_parent else None
    edges = U64()
    check(lib().B200_bfs(A.h, src, max_level, level.ctypes.data, None if parent is None else parent.ctypes.data, 0,
                         C.byref(edges)))
    return level, parent, edges.value


def get_stat(name):
    return lib().B200_get_stat(name.encode())


def reset_stats():
    lib().B200_reset_stats()


def set_option(name, value):
    check(lib().B200_set_option(name.encode(), int(value)))


def sync():
    check(lib().B200_sync())


def wait_ticket(ticket):
    """B200_Ticket_wait: the buffer given to export_bitmap_async is complete when this returns."""
    check(lib().B200_Ticket_wait(C.byref(ticket)))


OUT_AUTO, OUT_BITMAP, OUT_CSR = 0, 1, 2


def traverse_batch(sources, hop_matrices, fmt=OUT_AUTO, out_bitmap=None, out_p=None, out_j=None):
    """B200_traverse_batch: the batched traversal an operator runs (cond_traverse.rs:600-608), host to host, as ONE C call.
    F(i, sources[i]) = 1, F <- F * hop_matrices[0] * ... ; the result lands in `out_bitmap` (uint64[nsrc, ceil(n/64)], pinned for
    overlapped copies) and / or (`out_p` uint64[nsrc+1], `out_j` uint32[capacity]).  Returns (flops, nvals or None, format)."""
    sources = _u64arr(sources)
    hs = (C.c_void_p * len(hop_matrices))(*[m.h for m in hop_matrices])
    nv, fl, chosen = U64(0), U64(0), C.c_int(0)
    wpr = out_bitmap.shape[1] if out_bitmap is not None else 0
    info = lib().B200_traverse_batch(sources.ctypes.data, len(sources), hs, len(hop_matrices), fmt,
                                     None if out_bitmap is None else out_bitmap.ctypes.data, wpr,
                                     None if out_p is None else out_p.ctypes.data, None if out_j is None else out_j.ctypes.data,
                                     0 if out_j is None else len(out_j), C.byref(nv), C.byref(fl), C.byref(chosen))
    if info == -103:                     # GrB_INSUFFICIENT_SPACE: nv holds the entries needed
        raise BufferError(f"out_j holds {0 if out_j is None else len(out_j)} entries, the result has {nv.value}")
    check(info)
    return fl.value, (None if nv.value == 2 ** 64 - 1 else nv.value), chosen.value


def traverse_to_host(sources, A, hops, out_bitmap, sub_batches=None):
    """Dense-result form of traverse_batch: `hops` times the same matrix, bitmap hand-off in 128-row slices whose device-to-host
    copies overlap the next slice's hops (the result transfer, not the GPU, bounds this call).  Returns the flops."""
    del sub_batches                      # the slicing lives in the library now
    return traverse_batch(sources, [A] * hops, OUT_BITMAP, out_bitmap=out_bitmap)[0]


def reach_batch(sources, A, max_hops=None, include_sources=False):
    """B200_reach_batch: `multi_source_reach` below as ONE C-ABI call (what a Rust operator would bind).  Returns (R, levels)."""
    sources = _u64arr(sources)
    h, lv = P(), I64()
    check(lib().B200_reach_batch(C.byref(h), sources.ctypes.data, len(sources), A.h, -1 if max_hops is None else int(max_hops),
                                 1 if include_sources else 0, C.byref(lv)))
    return Matrix(0, 0, bool, _handle=h), lv.value


def multi_source_reach(sources, A, max_hops=None, include_sources=False):
    """Reachability from up to 1024 sources at once by level-synchronous mxm -- the multiplicity-insensitive core of a
    variable-length traversal `-[*1..k]->` with `emit_path = false` (cond_var_len_traverse.rs:196) and of the
    BFS phase of allShortestPaths (all_shortest_paths.rs:7-25): row i of the result holds every vertex reachable from
    sources[i] by a walk of 1..max_hops edges (no bound: until no row finds a new vertex).
    Each level is C<!R, replace> = F*A (matrix.rs:1386: GrB_DESC_RSC, structural complement of the reached set) followed
    by R = R u F (matrix.rs:1398-1400) -- both stay in device frontier form.  Returns (R, levels run).
    With `include_sources` the zero-length walk (i, sources[i]) is part of R and never re-discovered."""
    sources = _u64arr(sources)
    nsrc, n = len(sources), A.ncols()
    rows = np.arange(nsrc, dtype=np.uint64)
    F = Matrix(nsrc, n, bool)
    F.build(rows, sources)
    R = Matrix(nsrc, n, bool)
    if include_sources:
        R.build(rows, sources)
    level = 0
    while max_hops is None or level < max_hops:
        if R.nvals():
            F.mxm(F, A, R, Descriptor.RSC)
        else:
            F.lmxm(A)
        if F.nvals() == 0:
            break
        R.element_wise_add(None, None, F, None)
        level += 1
    return R, level

