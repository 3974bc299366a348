"""Serialization boundary: GxB_Container round trips, payload-vector load / unload, vector blobs, and the
reference's decode-time validation (vector.rs:652-686 transcribed).  Host-resident matrices only, so this runs without a GPU."""
import ctypes as C

import numpy as np
import pytest

import falkordb_b200 as fb
from falkordb_b200 import serial
from falkordb_b200._lib import lib, check, P, U64, GrbError
from falkordb_b200.grb import Matrix


def make(nrows, ncols, entries, valued=False):
    m = Matrix(nrows, ncols, np.uint64 if valued else bool)
    for e in entries:
        if valued:
            m.set(e[0], e[1], e[2])
        else:
            m.set(e[0], e[1])
    m.wait()
    return m


def test_container_struct_layout_matches_the_reference_bindings():
    c = serial.Container
    assert C.sizeof(c) == 608
    want = {"nrows": 0, "ncols": 8, "nrows_nonempty": 16, "ncols_nonempty": 24, "nvals": 32, "u64_future": 40, "format": 128,
            "orientation": 132, "header_arena": 136, "u32_future": 140, "p": 192, "h": 200, "b": 208, "i": 216, "x": 224,
            "vector_future": 232, "Y": 320, "matrix_future": 328, "iso": 448, "jumbled": 449, "bool_future": 450, "void_future": 480}
    for f, off in want.items():                       # mod.rs:14193-14236
        assert getattr(c, f).offset == off, f


@pytest.mark.parametrize("valued", [False, True])
def test_matrix_encode_decode_round_trip(valued):
    ent = [(0, 3), (0, 70000), (2, 0), (5, 5), (5, 6), (99, 1)]
    if valued:
        ent = [(i, j, 10 * k + 7) for k, (i, j) in enumerate(ent)]
        ent[2] = (2, 0, 0)                              # an explicit zero value survives (tensor.rs:1427-1476: edge id 0)
    m = make(100, 80000, ent, valued)
    before = list(m.iter())
    stream = serial.encode_matrix(m.h)
    assert list(m.iter()) == before                     # encode leaves the matrix usable (matrix.rs:535-536 reload)
    assert m.nrows() == 100 and m.ncols() == 80000
    assert len(stream) == 1 + 5 * 5 and len(stream[0]) == 608
    hdr = serial.Container.from_buffer_copy(stream[0])
    assert (hdr.nrows, hdr.ncols, hdr.nvals, hdr.format, hdr.orientation, hdr.iso) == (100, 80000, 6, 2, 0, not valued)
    h2 = serial.decode_matrix(stream)
    m2 = Matrix(0, 0, np.uint64 if valued else bool, _handle=h2)
    assert (m2.nrows(), m2.ncols(), m2.nvals()) == (100, 80000, 6)
    assert list(m2.iter()) == before


def test_hypersparse_container_for_huge_dimensions():
    n = 1 << 60                                          # Tensor.me is 2^60 x 2^60 (tensor.rs:254)
    m = make(n, n, [(5, 1 << 40), ((1 << 59) + 3, 2), ((1 << 59) + 3, (1 << 59))])
    before = list(m.iter())
    stream = serial.encode_matrix(m.h)
    hdr = serial.Container.from_buffer_copy(stream[0])
    assert hdr.format == 1 and hdr.nrows_nonempty == 2 and hdr.nvals == 3          # GxB_HYPERSPARSE
    m2 = Matrix(0, 0, bool, _handle=serial.decode_matrix(stream))
    assert list(m2.iter()) == before and m2.nrows() == n


def test_empty_matrix_round_trip():
    m = Matrix(7, 9, bool)
    m2 = Matrix(0, 0, bool, _handle=serial.decode_matrix(serial.encode_matrix(m.h)))
    assert (m2.nrows(), m2.ncols(), m2.nvals()) == (7, 9, 0)


def _stream(valued=False):
    m = make(6, 6, [(0, 1), (0, 4), (3, 2), (5, 5)] if not valued else [(0, 1, 3), (3, 2, 9)], valued)
    return serial.encode_matrix(m.h)


def _vec(stream, k):
    """position of payload vector k (0 = x, 1 = h, 2 = p, 3 = i, 4 = b) inside the stream"""
    return 1 + 5 * k


@pytest.mark.parametrize("mutate,what", [
    (lambda s: s.__setitem__(0, s[0][:100]), "container buffer too small"),
    (lambda s: s.__setitem__(_vec(s, 2) + 3, s[_vec(s, 2) + 3] + 8), "does not match buffer length"),      # p: n_bytes lie
    (lambda s: s.__setitem__(_vec(s, 3) + 1, b"GrB_UINT32"), "NUL-terminated"),                              # i: no NUL
    (lambda s: s.__setitem__(_vec(s, 3) + 1, b"Gr\0B\0"), "NUL-terminated"),                                 # interior NUL
    (lambda s: s.__setitem__(_vec(s, 3) + 1, b"no_such_type\0"), "GxB_Type_from_name failed"),
])
def test_decode_rejects_malformed_streams_like_the_reference(mutate, what):
    s = _stream()
    mutate(s)
    with pytest.raises(ValueError, match=what):
        serial.decode_matrix(s)


def _patch_u(stream, k, index, value, width):
    pos = _vec(stream, k)
    a = bytearray(stream[pos])
    a[index * width:(index + 1) * width] = int(value).to_bytes(width, "little")
    stream[pos] = bytes(a)


@pytest.mark.parametrize("mutate", [
    lambda s: _patch_u(s, 2, 1, 99, 8),          # p not monotone / beyond nvals
    lambda s: _patch_u(s, 2, 0, 1, 8),           # p[0] != 0
    lambda s: _patch_u(s, 3, 0, 6, 4),           # column index == ncols
    lambda s: _patch_u(s, 3, 1, 1, 4),           # duplicate column inside a row
    lambda s: s.__setitem__(0, s[0][:32] + (77).to_bytes(8, "little") + s[0][40:]),   # nvals disagrees with p
    lambda s: s.__setitem__(0, s[0][:128] + (4).to_bytes(4, "little") + s[0][132:]),  # bitmap format: unsupported
])
def test_load_validates_untrusted_container_payloads(mutate):
    s = _stream()
    mutate(s)
    with pytest.raises(GrbError):
        serial.decode_matrix(s)


def test_vector_blob_round_trip_and_rejections():
    L = lib()
    serial._sig()
    v = P()
    n = (1 << 60)
    check(L.GrB_Vector_new(C.byref(v), C.c_void_p.in_dll(L, "GrB_BOOL"), n))
    ids = [0, 7, 123456789012, (1 << 60) - 1]            # edge ids are the vector's indices (vector.rs:150-153)
    for i in ids:
        check(L.GrB_Vector_setElement_BOOL(v, True, i))
    blob = serial.vector_to_blob(v)
    w = serial.vector_from_blob(blob)
    sz, nv = U64(), U64()
    check(L.GrB_Vector_size(C.byref(sz), w)); check(L.GrB_Vector_nvals(C.byref(nv), w))
    assert sz.value == n and nv.value == len(ids)
    I = (U64 * len(ids))(); X = (C.c_bool * len(ids))(); cap = U64(len(ids))
    check(L.GrB_Vector_extractTuples_BOOL(I, X, C.byref(cap), w))
    assert list(I) == ids and all(X)
    for bad in (blob[:10], blob[:-1], blob + b"\0", b"XXXX" + blob[4:], blob[:24] + (99).to_bytes(8, "little") + blob[32:]):
        with pytest.raises(GrbError):
            serial.vector_from_blob(bad)
    L.GrB_Vector_free(C.byref(v)); L.GrB_Vector_free(C.byref(w))


def test_type_names_round_trip():
    L = lib()
    serial._sig()
    for sym in ("GrB_BOOL", "GrB_UINT64", "GrB_UINT32", "GrB_INT64"):
        t = C.c_void_p.in_dll(L, sym)
        name = C.create_string_buffer(128)
        check(L.GrB_Type_get_String(t, name, serial.GrB_NAME))
        assert name.value == sym.encode()
        back = P()
        check(L.GxB_Type_from_name(C.byref(back), name.value))
        assert back.value == t.value
    back = P()
    check(L.GxB_Type_from_name(C.byref(back), b"bool"))       # the C name form used by the reference's mock (vector.rs:662)
    assert back.value == C.c_void_p.in_dll(L, "GrB_BOOL").value
