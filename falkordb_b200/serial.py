"""Host-side mirror of the reference's RDB encode / decode of a GraphBLAS matrix and of its payload vectors
(graph/src/graph/graphblas/matrix.rs:428-546, vector.rs:150-420) over the C ABI's serialization entry points
(GxB_Container_*, GxB_unload_Matrix_into_Container / GxB_load_Matrix_from_Container, GxB_Vector_load / unload,
GxB_Vector_serialize / deserialize).  A "stream" here is the list of items the reference's Writer receives:
bytes objects (write_buffer) and ints (write_unsigned / write_signed)."""
import ctypes as C

from ._lib import lib, check, P, U64, I64

CONTAINER_STRUCT_SIZE = 608          # mod.rs:14191
GxB_MAX_NAME_LEN = 128
GrB_NAME = 10
_libc = C.CDLL(None)
_libc.malloc.restype = C.c_void_p
_libc.malloc.argtypes = [C.c_size_t]
_libc.free.argtypes = [C.c_void_p]


class Container(C.Structure):
    """mod.rs:14165-14188"""
    _fields_ = [("nrows", U64), ("ncols", U64), ("nrows_nonempty", I64), ("ncols_nonempty", I64), ("nvals", U64),
                ("u64_future", U64 * 11), ("format", C.c_int32), ("orientation", C.c_int32), ("header_arena", C.c_int32),
                ("u32_future", C.c_uint32 * 13), ("p", P), ("h", P), ("b", P), ("i", P), ("x", P), ("vector_future", P * 11),
                ("Y", P), ("matrix_future", P * 15), ("iso", C.c_bool), ("jumbled", C.c_bool), ("bool_future", C.c_bool * 30),
                ("void_future", P * 16)]


assert C.sizeof(Container) == CONTAINER_STRUCT_SIZE


def _sig():
    L = lib()
    if getattr(L, "_serial_sig", False):
        return L
    CP = C.POINTER(Container)
    L.GxB_Container_new.argtypes = [C.POINTER(CP)]
    L.GxB_Container_free.argtypes = [C.POINTER(CP)]
    L.GxB_unload_Matrix_into_Container.argtypes = [P, CP, P]
    L.GxB_load_Matrix_from_Container.argtypes = [P, CP, P]
    L.GxB_Vector_unload.argtypes = [P, C.POINTER(P), C.POINTER(P), C.POINTER(U64), C.POINTER(U64), C.POINTER(C.c_int), P]
    L.GxB_Vector_load.argtypes = [P, C.POINTER(P), P, U64, U64, C.c_int, P]
    L.GxB_Vector_serialize.argtypes = [C.POINTER(P), C.POINTER(U64), P, P]
    L.GxB_Vector_deserialize.argtypes = [C.POINTER(P), P, C.c_char_p, U64, P]
    L.GrB_Type_get_String.argtypes = [P, C.c_char_p, C.c_int]
    L.GxB_Type_from_name.argtypes = [C.POINTER(P), C.c_char_p]
    for n in ("GxB_Container_new", "GxB_Container_free", "GxB_unload_Matrix_into_Container", "GxB_load_Matrix_from_Container",
              "GxB_Vector_unload", "GxB_Vector_load", "GxB_Vector_serialize", "GxB_Vector_deserialize", "GrB_Type_get_String",
              "GxB_Type_from_name"):
        getattr(L, n).restype = C.c_int
    L._serial_sig = True
    return L


def encode_vector(vh, out):
    """<Vector<bool> as Encode>::encode, vector.rs:241-309: unload, write (array, type name + NUL, n, bytes, handling), reload"""
    L = _sig()
    arr, typ, n, nbytes, handling = P(), P(), U64(), U64(), C.c_int()
    check(L.GxB_Vector_unload(vh, C.byref(arr), C.byref(typ), C.byref(n), C.byref(nbytes), C.byref(handling), None))
    name = C.create_string_buffer(GxB_MAX_NAME_LEN)
    check(L.GrB_Type_get_String(typ, name, GrB_NAME))
    out.append(C.string_at(arr, nbytes.value) if nbytes.value else b"")
    out.append(name.value + b"\0")
    out += [n.value, nbytes.value, handling.value]
    check(L.GxB_Vector_load(vh, C.byref(arr), typ, n, nbytes, handling, None))


def decode_vector(stream):
    """<Vector<bool> as Decode>::decode, vector.rs:311-413, including its validation of the untrusted payload"""
    L = _sig()
    arr_data, type_name, n_entries, n_bytes, handling = (stream.pop(0) for _ in range(5))
    if n_bytes != len(arr_data):
        raise ValueError(f"Vector decode: declared byte length {n_bytes} does not match buffer length {len(arr_data)}")
    if not type_name or type_name[-1] != 0 or 0 in type_name[:-1]:
        raise ValueError("Vector decode: type name is not NUL-terminated")
    typ = P()
    info = L.GxB_Type_from_name(C.byref(typ), type_name)
    if info != 0:
        raise ValueError(f"Vector decode: GxB_Type_from_name failed: {info}")
    v = P()
    check(L.GrB_Vector_new(C.byref(v), typ, 0))
    ptr = P(_libc.malloc(n_bytes)) if n_bytes else P()
    if n_bytes:
        C.memmove(ptr, arr_data, n_bytes)
    info = L.GxB_Vector_load(v, C.byref(ptr), typ, n_entries, n_bytes, handling, None)
    if info != 0:
        if ptr:
            _libc.free(ptr)
        L.GrB_Vector_free(C.byref(v))
        raise ValueError(f"Vector decode: GxB_Vector_load failed: {info}")
    return v


def encode_matrix(mh):
    """<Matrix<T> as Encode>::encode, matrix.rs:508-546.  Returns the stream; the matrix is usable again afterwards."""
    L = _sig()
    c = C.POINTER(Container)()
    check(L.GxB_Container_new(C.byref(c)))
    out = []
    try:
        check(L.GxB_unload_Matrix_into_Container(mh, c, None))
        out.append(C.string_at(c, CONTAINER_STRUCT_SIZE))
        for f in ("x", "h", "p", "i", "b"):
            encode_vector(getattr(c.contents, f), out)
        check(L.GxB_load_Matrix_from_Container(mh, c, None))
    finally:
        L.GxB_Container_free(C.byref(c))
    return out


def decode_matrix(stream):
    """<Matrix<T> as Decode>::decode, matrix.rs:428-506.  Returns a new GrB_Matrix handle."""
    L = _sig()
    stream = list(stream)
    blob = stream.pop(0)
    if len(blob) < CONTAINER_STRUCT_SIZE:
        raise ValueError(f"container buffer too small: {len(blob)} bytes < {CONTAINER_STRUCT_SIZE} bytes required")
    c = C.POINTER(Container)()
    check(L.GxB_Container_new(C.byref(c)))
    own = {f: getattr(c.contents, f) for f in ("x", "h", "p", "i", "b")}   # the empty vectors Container_new made
    try:
        C.memmove(c, blob, CONTAINER_STRUCT_SIZE)
        for f in ("x", "h", "b", "i", "p", "Y"):
            setattr(c.contents, f, None)
        for f, v in own.items():                 # the reference leaks nothing either: it frees via Container_free below
            vv = P(v)
            L.GrB_Vector_free(C.byref(vv))
        for f in ("x", "h", "p", "i", "b"):
            setattr(c.contents, f, decode_vector(stream))
        m = P()
        check(L.GrB_Matrix_new(C.byref(m), C.c_void_p.in_dll(L, "GrB_BOOL"), 0, 0))
        info = L.GxB_load_Matrix_from_Container(m, c, None)
        if info != 0:
            L.GrB_Matrix_free(C.byref(m))
            check(info)
        check(L.GrB_Matrix_wait(m, 1))
        return m
    finally:
        L.GxB_Container_free(C.byref(c))


def vector_to_blob(vh):
    """Vector::encode_blob, vector.rs:157-174"""
    L = _sig()
    blob, size = P(), U64()
    check(L.GxB_Vector_serialize(C.byref(blob), C.byref(size), vh, None))
    data = C.string_at(blob, size.value)
    _libc.free(blob)
    return data


def vector_from_blob(data):
    """Vector::decode_blob, vector.rs:177-195"""
    L = _sig()
    v = P()
    check(L.GxB_Vector_deserialize(C.byref(v), None, data, len(data), None))
    return v
